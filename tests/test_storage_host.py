"""CPU-only host logic of the drop-in RolloutStorage: state machine, shapes, dtypes, asserts."""
import pytest
import torch

from a2c_ppo_acktr.storage import RolloutStorage


class Discrete(object):
    def __init__(self, n):
        self.n = n


class Box(object):
    def __init__(self, d):
        self.shape = (d,)


def test_shapes_and_dtypes():
    st = RolloutStorage(5, 3, (4, 8, 8), Discrete(6), 1)
    assert st.obs.shape == (6, 3, 4, 8, 8) and st.obs.dtype == torch.float32
    assert st.actions.shape == (5, 3, 1) and st.actions.dtype == torch.int64
    assert st.masks.shape == (6, 3, 1) and bool((st.masks == 1).all()) and bool((st.bad_masks == 1).all())
    assert st.rewards.shape == (5, 3, 1) and st.returns.shape == (6, 3, 1)
    st2 = RolloutStorage(5, 3, (11,), Box(3), 64)
    assert st2.actions.shape == (5, 3, 3) and st2.actions.dtype == torch.float32
    assert st2.recurrent_hidden_states.shape == (6, 3, 64)
    st3 = RolloutStorage(2, 2, (4, 8, 8), Discrete(6), 1, obs_dtype=torch.uint8)
    assert st3.obs.dtype == torch.uint8


def test_insert_after_update_state_machine():
    T, N = 3, 2
    st = RolloutStorage(T, N, (2,), Discrete(4), 1)
    for t in range(T):
        st.insert(torch.full((N, 2), float(t + 1)), torch.zeros(N, 1), torch.full((N, 1), t, dtype=torch.int64),
                  torch.full((N, 1), -float(t)), torch.full((N, 1), 10.0 + t), torch.full((N, 1), 0.5 * t),
                  torch.full((N, 1), float(t % 2)), torch.ones(N, 1))
    assert st.step == 0                                   # wrapped (storage.py:58)
    assert st.obs[:, 0, 0].tolist() == [0.0, 1.0, 2.0, 3.0]
    assert st.actions[:, 0, 0].tolist() == [0, 1, 2]
    assert st.masks[:, 0, 0].tolist() == [1.0, 0.0, 1.0, 0.0]
    st.after_update()
    assert st.obs[0, 0, 0].item() == 3.0 and st.masks[0, 0, 0].item() == 0.0


def test_too_few_rows_assertion_messages():
    st = RolloutStorage(2, 2, (2,), Discrete(4), 1)
    with pytest.raises(AssertionError, match="PPO requires the number of processes"):
        next(st.feed_forward_generator(None, num_mini_batch=8))
    with pytest.raises(AssertionError, match="to be greater than or equal to the number of"):
        next(st.recurrent_generator(torch.zeros(2, 2, 1), 4))


def test_frame_ring_reproduces_the_reference_frame_stack():
    """RolloutStorage(frame_ring=True) stores every 84 x 84 frame once; reading `obs[t]` must give exactly the rolling
    4-stack the reference's VecPyTorchFrameStack builds (envs.py:218-259: shift left, CLEAR the stack of an environment
    whose episode ended, write the newest frame), across `after_update` boundaries."""
    from a2c_ppo_acktr.storage import RolloutStorage

    class Discrete(object):
        n = 6
    T, N, C = 7, 5, 4
    g = torch.Generator().manual_seed(3)
    stacked = RolloutStorage(T, N, (C, 84, 84), Discrete(), 1, obs_dtype=torch.uint8)
    ring = RolloutStorage(T, N, (C, 84, 84), Discrete(), 1, obs_dtype=torch.uint8, frame_ring=True)
    assert ring.frames.numel() * 4 < stacked.obs.numel() * 1.5          # ~ a quarter of the bytes
    stack = torch.zeros(N, C, 84, 84, dtype=torch.uint8)               # the wrapper's stacked_obs
    stack[:, -1] = torch.randint(0, 256, (N, 84, 84), generator=g, dtype=torch.uint8)     # reset()
    stacked.obs[0].copy_(stack)
    ring.obs[0].copy_(stack)                                            # reference main.py:97
    assert torch.equal(ring.obs[0], stack)
    z = torch.zeros(N, 1)
    for it in range(3):
        for t in range(T):
            done = torch.rand(N, generator=g) < 0.25
            frame = torch.randint(0, 256, (N, 84, 84), generator=g, dtype=torch.uint8)
            stack[:, :-1] = stack[:, 1:].clone()
            stack[done] = 0
            stack[:, -1] = frame
            masks = (~done).float().view(N, 1)
            for st in (stacked, ring):
                st.insert(stack, z, torch.zeros(N, 1, dtype=torch.int64), z, z, z, masks, torch.ones(N, 1))
        assert tuple(ring.obs.shape) == tuple(stacked.obs.shape)
        for t in range(T + 1):
            assert torch.equal(ring.obs[t], stacked.obs[t]), (it, t)
        assert torch.equal(ring.obs[:-1], stacked.obs[:-1])
        stacked.after_update()
        ring.after_update()
        assert torch.equal(ring.obs[0], stacked.obs[0])
