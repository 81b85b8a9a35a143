"""GPU parity of the GAIL discriminator (csrc/gail.cu, algo/gail.py) against the golden produced by the unmodified
reference (a2c_ppo_acktr/algo/gail.py:11-110) and against the CPU oracle.  Tolerance: north-star 1e-4 (fp32)."""
import numpy as np
import pytest
import torch

import helpers
from oracle import gail as o_gail

pytestmark = pytest.mark.gpu

from a2c_ppo_acktr.algo import gail  # noqa: E402
from a2c_ppo_acktr.storage import RolloutStorage  # noqa: E402

DEV = "cuda:0"


class Box(object):
    def __init__(self, d):
        self.shape = (d,)


def obsfilt(x, update=False):
    return np.clip((x - 0.3) / np.sqrt(4.0 + 1e-8), -10.0, 10.0)


def _storage(obs, actions, masks):
    T, N = actions.shape[:2]
    st = RolloutStorage(T, N, (obs.shape[-1],), Box(actions.shape[-1]), 1)
    st.obs.copy_(obs)
    st.actions.copy_(actions)
    st.masks.copy_(masks)
    st.to(DEV)
    return st


def test_discriminator_vs_reference_golden():
    g = helpers.load("gail")
    obs, actions, masks = (torch.from_numpy(g["in_" + k]) for k in ("obs", "actions", "masks"))
    T, N = actions.shape[:2]
    torch.manual_seed(1)
    disc = gail.Discriminator(obs.shape[-1] + actions.shape[-1], 100, torch.device(DEV))
    for k, v in disc.state_dict().items():
        assert np.array_equal(v.cpu().numpy(), g["init_" + k]), k              # reference keys and init stream
    st = _storage(obs, actions, masks)
    ds = torch.utils.data.TensorDataset(torch.from_numpy(g["in_expert_states"]), torch.from_numpy(g["in_expert_actions"]))
    loader = torch.utils.data.DataLoader(ds, batch_size=16, shuffle=True, drop_last=True)
    torch.manual_seed(7)
    losses = [disc.update(loader, st, obsfilt) for _ in range(2)]
    np.testing.assert_allclose(losses, g["losses"], rtol=1e-4)
    num = den = 0.0
    for k, v in disc.state_dict().items():
        want = g["final_" + k].astype(np.float64)
        num += float(((v.cpu().numpy() - want) ** 2).sum())
        den += float((want ** 2).sum())
        np.testing.assert_allclose(v.cpu().numpy(), g["final_" + k], rtol=2e-3, atol=2e-5, err_msg=k)
    assert (num / den) ** 0.5 <= 1e-4
    # the per-step loop of reference main.py:152-155 ...
    rewards = torch.stack([disc.predict_reward(st.obs[t], st.actions[t], 0.99, st.masks[t]) for t in range(T)])
    np.testing.assert_allclose(rewards.cpu().numpy(), g["rewards"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose([disc.ret_rms.mean[0], disc.ret_rms.var[0], disc.ret_rms.count], g["rms"], rtol=1e-5)
    np.testing.assert_allclose(disc.returns.cpu().numpy(), g["returns_state"], rtol=1e-4, atol=1e-5)
    frozen = disc.predict_reward(st.obs[0], st.actions[0], 0.99, st.masks[0], update_rms=False)
    np.testing.assert_allclose(frozen.cpu().numpy(), g["rewards_frozen"], rtol=1e-4, atol=1e-5)


def test_batched_reward_pass_equals_the_per_step_loop():
    """predict_rewards(rollouts, gamma): two launches over [T, N] must equal T calls of predict_reward exactly."""
    g = torch.Generator().manual_seed(3)
    T, N, Ds, Da = 64, 24, 17, 6
    obs, actions = torch.randn(T + 1, N, Ds, generator=g), torch.randn(T, N, Da, generator=g)
    masks = (torch.rand(T + 1, N, 1, generator=g) > 0.05).float()
    out = []
    for batched in (False, True):
        torch.manual_seed(1)
        disc = gail.Discriminator(Ds + Da, 100, torch.device(DEV))
        st = _storage(obs, actions, masks)
        if batched:
            disc.predict_rewards(st, 0.99)
            r = st.rewards.clone()
        else:
            r = torch.stack([disc.predict_reward(st.obs[t], st.actions[t], 0.99, st.masks[t]) for t in range(T)])
        out.append((r, disc.ret_rms.state.clone(), disc.returns.clone()))
    for a, b in zip(out[0], out[1]):
        assert torch.equal(a, b)


@pytest.mark.parametrize("B,Ds,Da,H", [(128, 11, 3, 100), (20, 5, 2, 33), (64, 100, 20, 128)])
def test_discriminator_step_vs_oracle(B, Ds, Da, H):
    """One update (several minibatch steps) on random data against the autograd oracle (double backward through the
    gradient penalty), incl. a batch size that is not a multiple of the 16-row CTA tile."""
    g = torch.Generator().manual_seed(9)
    T, N = 4 * B // 8 + 3, 8
    obs, actions = torch.randn(T + 1, N, Ds, generator=g), torch.randn(T, N, Da, generator=g)
    masks = torch.ones(T + 1, N, 1)
    es, ea = torch.randn(3 * B, Ds, generator=g), torch.randn(3 * B, Da, generator=g)
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(es, ea), batch_size=B, shuffle=True, drop_last=True)
    torch.manual_seed(1)
    disc = gail.Discriminator(Ds + Da, H, torch.device(DEV))
    p = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in disc.state_dict().items()}
    opt = torch.optim.Adam(list(p.values()))
    st = _storage(obs, actions, masks)
    torch.manual_seed(7)
    want = o_gail.update(p, opt, loader, obs[:-1].reshape(T * N, -1), actions.reshape(T * N, -1), lambda x, update=False: x)
    torch.manual_seed(7)
    got = disc.update(loader, st, lambda x, update=False: x)
    np.testing.assert_allclose(got, want, rtol=1e-4)
    num = den = 0.0
    for k, v in disc.state_dict().items():
        w = p[k].detach().numpy().astype(np.float64)
        num += float(((v.cpu().numpy() - w) ** 2).sum())
        den += float((w ** 2).sum())
    assert (num / den) ** 0.5 <= 1e-4
