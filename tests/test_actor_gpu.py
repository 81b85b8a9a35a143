"""GPU: the actor step -- sampling head (csrc/actor.cu) against oracle/actor.py on the same Philox stream, and the
fused `Policy.act_into(rollouts, step)` against `Policy.act` + the reference's insert semantics
(reference model.py:54-66, main.py:115-134, storage.py:46-58)."""
import numpy as np
import pytest
import torch

from oracle import actor

pytestmark = pytest.mark.gpu

from a2c_ppo_acktr.model import Policy  # noqa: E402
from a2c_ppo_acktr.storage import RolloutStorage  # noqa: E402

DEV = "cuda:0"


class Discrete(object):
    def __init__(self, n):
        self.n = n


class Box(object):
    def __init__(self, d):
        self.shape = (d,)


def _set_rng(rt, seed, offset):
    rt.rng_state().copy_(torch.tensor([seed, offset, 0, 0], dtype=torch.int64))


@pytest.mark.parametrize("B", [1, 77, 4096])
def test_categorical_head_vs_oracle(B):
    torch.manual_seed(0)
    pol = Policy((11,), Discrete(6)).to(DEV)
    rt = pol.runtime()
    z = (torch.randn(B, 7) * 2.0).to(DEV)
    act = torch.zeros(B, 1, dtype=torch.int64, device=DEV)
    val, lp = torch.zeros(B, 1, device=DEV), torch.zeros(B, 1, device=DEV)
    seed = 0x1234567890ABCDE
    for offset in (0, 1):                                     # the second launch must see the advanced stream
        if offset == 0:
            _set_rng(rt, seed, 0)
        rt.sample(z, act, val, lp, False)
        v, a, logp, u, cdf = actor.categorical(z.cpu().numpy(), seed, offset)
        got = act.view(-1).cpu().numpy()
        near = (np.abs(u[:, None] - cdf).min(1) < 1e-5)       # u on a CDF boundary: expf rounding may differ
        assert np.array_equal(got[~near], a[~near]) and near.mean() < 0.01
        assert torch.equal(val.view(-1).cpu(), torch.from_numpy(v))
        same = got == a
        np.testing.assert_allclose(lp.view(-1).cpu().numpy()[same], logp[same], rtol=1e-5, atol=1e-6)
    assert rt.rng_state().cpu().tolist() == [seed, 2, 0, 0]
    rt.sample(z, act, val, lp, True)
    assert torch.equal(act.view(-1), z[:, 1:].argmax(1))
    assert rt.rng_state().cpu().tolist() == [seed, 2, 0, 0]     # the mode draws nothing


def test_categorical_head_frequencies():
    torch.manual_seed(0)
    pol = Policy((11,), Discrete(6)).to(DEV)
    rt = pol.runtime()
    z = torch.randn(1, 7).repeat(200000, 1).to(DEV)
    act = torch.zeros(200000, 1, dtype=torch.int64, device=DEV)
    val, lp = torch.zeros(200000, 1, device=DEV), torch.zeros(200000, 1, device=DEV)
    rt.sample(z, act, val, lp, False)
    freq = torch.bincount(act.view(-1), minlength=6).double().cpu() / 200000.0
    p = torch.softmax(z[0, 1:].double().cpu(), 0)
    assert (freq - p).abs().max() < 0.005                     # ~4.5 sigma of a 0.5 / 200000 binomial


@pytest.mark.parametrize("A", [1, 3, 6])
def test_gaussian_head_vs_oracle(A):
    torch.manual_seed(0)
    pol = Policy((11,), Box(A)).to(DEV)
    with torch.no_grad():
        pol.dist.logstd._bias.copy_(torch.linspace(-0.5, 0.3, A).view(A, 1))
    rt = pol.runtime()
    B = 513
    z = torch.randn(B, 1 + A).to(DEV)
    act, val, lp = torch.zeros(B, A, device=DEV), torch.zeros(B, 1, device=DEV), torch.zeros(B, 1, device=DEV)
    _set_rng(rt, 77, 5)
    rt.sample(z, act, val, lp, False)
    logstd = pol.dist.logstd._bias.detach().view(-1).cpu().numpy()
    v, a, logp = actor.gaussian(z.cpu().numpy(), logstd, 77, 5)
    np.testing.assert_allclose(act.cpu().numpy(), a, rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(lp.view(-1).cpu().numpy(), logp, rtol=1e-3, atol=1e-3)
    # log-prob of the action actually written, exactly as evaluate_actions would compute it
    d = torch.distributions.Normal(z[:, 1:], torch.from_numpy(np.exp(logstd)).to(DEV))
    np.testing.assert_allclose(lp.view(-1).cpu().numpy(), d.log_prob(act).sum(-1).cpu().numpy(), rtol=1e-5, atol=1e-5)
    rt.sample(z, act, val, lp, True)
    assert torch.equal(act, z[:, 1:])


@pytest.mark.parametrize("kind", ["cnn", "cnn_gru", "mlp_box", "mlp_gru"])
def test_act_into_equals_act_and_fills_the_slots(kind):
    torch.manual_seed(1)
    cnn = kind.startswith("cnn")
    recurrent = kind.endswith("gru")
    obs_shape = (4, 84, 84) if cnn else (11,)
    space = Discrete(6) if kind != "mlp_box" else Box(3)
    pol = Policy(obs_shape, space, base_kwargs={"recurrent": recurrent}).to(DEV)
    T, N, H = 4, 8, pol.recurrent_hidden_state_size
    st = RolloutStorage(T, N, obs_shape, space, H, obs_dtype=torch.uint8 if cnn else torch.float32)
    st.to(DEV)
    if cnn:
        st.obs.random_(0, 256)
    else:
        st.obs.normal_()
    st.recurrent_hidden_states[0].normal_()
    st.masks[1, 2] = 0.0
    rt = pol.runtime()
    for det in (True, False):
        for t in range(T):
            _set_rng(rt, 4242, 10 * t)
            v0, a0, lp0, h0 = pol.act(st.obs[t], st.recurrent_hidden_states[t], st.masks[t], deterministic=det)
            _set_rng(rt, 4242, 10 * t)
            v1, a1, lp1, h1 = pol.act_into(st, t, deterministic=det)
            assert torch.equal(v0, v1) and torch.equal(a0, a1) and torch.equal(lp0, lp1) and torch.equal(h0, h1)
            assert v1.data_ptr() == st.value_preds[t].data_ptr() and h1.data_ptr() == st.recurrent_hidden_states[t + 1].data_ptr()
            assert torch.equal(st.actions[t], a0) and torch.equal(st.action_log_probs[t], lp0)
            # log-prob consistent with evaluate_actions on the same inputs (what PPO's ratio starts from)
            _, lpe, _, _ = pol.evaluate_actions(st.obs[t], st.recurrent_hidden_states[t], st.masks[t], a0)
            np.testing.assert_allclose(lpe.detach().cpu().numpy(), lp0.cpu().numpy(), rtol=1e-5, atol=1e-6)
            # the reference's insert() with the returned views is a no-op on these slots
            before = (st.value_preds[t].clone(), st.actions[t].clone(), st.recurrent_hidden_states[t + 1].clone())
            st.step = t
            st.insert(st.obs[t + 1].clone(), h1, a1, lp1, v1, torch.zeros(N, 1, device=DEV), st.masks[t + 1].clone(),
                      st.bad_masks[t + 1].clone())
            assert torch.equal(before[0], st.value_preds[t]) and torch.equal(before[1], st.actions[t])
            assert torch.equal(before[2], st.recurrent_hidden_states[t + 1])
    assert sum(1 for k in rt.cache if k[0] == "act") == 1


def test_act_into_graph_replay_matches_eager_and_keeps_sampling():
    """Third and later uses of a slot replay a CUDA graph: same results as the eager launch for the same stream
    state, and consecutive replays keep drawing fresh actions (the stream advances on the device)."""
    torch.manual_seed(1)
    pol = Policy((4, 84, 84), Discrete(6), base_kwargs={"recurrent": True}).to(DEV)
    T, N = 2, 16
    st = RolloutStorage(T, N, (4, 84, 84), Discrete(6), 512, obs_dtype=torch.uint8)
    st.to(DEV)
    st.obs.random_(0, 256)
    rt = pol.runtime()
    outs = []
    for rep in range(4):                                      # eager, capture + replay, replay, replay
        _set_rng(rt, 99, 0)
        v, a, lp, h = pol.act_into(st, 0)
        outs.append((v.clone(), a.clone(), lp.clone(), h.clone()))
    plan = rt.act_plan(N, torch.uint8)
    assert any(isinstance(g, tuple) for g in plan.graphs.values())
    for o in outs[1:]:
        for x, y in zip(outs[0], o):
            assert torch.equal(x, y)
    draws = []
    for rep in range(6):
        draws.append(pol.act_into(st, 0)[1].clone())
    assert rt.rng_state().cpu().tolist()[1] == 7
    assert any(not torch.equal(draws[0], d) for d in draws[1:])


def test_sampling_stream_is_sharded_persistent_and_reseedable(tmp_path):
    """identically seeded ranks must not draw identical noise for their local environment i (the Philox
    counter carries the GLOBAL environment index), the stream must survive pickling / device moves / re-flattening,
    and a later torch.manual_seed must take effect."""
    torch.manual_seed(5)
    pol = Policy((11,), Discrete(6)).to(DEV)
    rt = pol.runtime()
    B = 64
    z = (torch.randn(B, 7) * 2.0).to(DEV)
    act = torch.zeros(B, 1, dtype=torch.int64, device=DEV)
    val, lp = torch.zeros(B, 1, device=DEV), torch.zeros(B, 1, device=DEV)
    seed = int(pol.sampling_state()[0].item())
    from a2c_ppo_acktr import _engine
    assert seed == _engine.derive_sampling_seed(5)
    rt.sample(z, act, val, lp, False)
    a_rank0 = act.clone()
    # "rank 1": same seed, same offset, environments 64..127
    pol.sampling_state()[1] = 0
    pol.set_sampling_shard(64)
    rt.sample(z, act, val, lp, False)
    v, a, logp, u, cdf = actor.categorical(z.cpu().numpy(), seed, 0, row_base=64)
    near = (np.abs(u[:, None] - cdf).min(1) < 1e-5)
    assert np.array_equal(act.view(-1).cpu().numpy()[~near], a[~near])
    assert not torch.equal(act, a_rank0)
    assert pol.sampling_state().cpu().tolist() == [seed, 1, 0, 64]
    # survives re-flattening (a new runtime) and pickling
    pol._flatten()
    assert pol.runtime() is not rt and pol.runtime().rng_state().cpu().tolist() == [seed, 1, 0, 64]
    path = tmp_path / "p.pt"
    torch.save([pol, None], path)
    pol2, _ = torch.load(path, weights_only=False)
    assert pol2.runtime().rng_state().cpu().tolist() == [seed, 1, 0, 64]
    # re-seeding restarts the stream under the new key, keeping the shard offset
    torch.manual_seed(6)
    obs = torch.randn(4, 11, device=DEV)
    pol.act(obs, torch.zeros(4, 1, device=DEV), torch.ones(4, 1, device=DEV))
    assert pol.sampling_state().cpu().tolist() == [_engine.derive_sampling_seed(6), 1, 0, 64]
