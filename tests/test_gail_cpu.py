"""CPU: the GAIL oracle (oracle/gail.py) against the golden produced by the unmodified reference
(tests/golden/make_golden.py gail -> reference a2c_ppo_acktr/algo/gail.py:11-110)."""
import numpy as np
import torch

import helpers
from oracle import gail as o_gail


def obsfilt(x, update=False):
    return np.clip((x - 0.3) / np.sqrt(4.0 + 1e-8), -10.0, 10.0)


def test_gail_oracle_matches_reference_golden():
    g = helpers.load("gail")
    obs, actions, masks = (torch.from_numpy(g["in_" + k]) for k in ("obs", "actions", "masks"))
    T, N = actions.shape[:2]
    p0 = o_gail.init_params(obs.shape[-1] + actions.shape[-1], 100, seed=1)
    for k, v in p0.items():
        assert np.array_equal(v.numpy(), g["init_" + k]), k                    # same constructor RNG stream
    p = {k: v.clone().requires_grad_(True) for k, v in p0.items()}
    opt = torch.optim.Adam(list(p.values()))
    ds = torch.utils.data.TensorDataset(torch.from_numpy(g["in_expert_states"]), torch.from_numpy(g["in_expert_actions"]))
    loader = torch.utils.data.DataLoader(ds, batch_size=16, shuffle=True, drop_last=True)
    obs_flat, act_flat = obs[:-1].reshape(T * N, -1), actions.reshape(T * N, -1)
    torch.manual_seed(7)
    losses = [o_gail.update(p, opt, loader, obs_flat, act_flat, obsfilt) for _ in range(2)]
    np.testing.assert_allclose(losses, g["losses"], rtol=1e-5)
    for k, v in p.items():
        np.testing.assert_allclose(v.detach().numpy(), g["final_" + k], rtol=1e-4, atol=1e-6, err_msg=k)
    rs = o_gail.RewardState()
    pd = {k: v.detach() for k, v in p.items()}
    rewards = torch.stack([o_gail.predict_reward(pd, rs, obs[t], actions[t], 0.99, masks[t]) for t in range(T)])
    np.testing.assert_allclose(rewards.numpy(), g["rewards"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose([float(np.asarray(rs.ret_rms.mean).reshape(-1)[0]), float(np.asarray(rs.ret_rms.var).reshape(-1)[0]),
                                rs.ret_rms.count], g["rms"], rtol=1e-6)
    frozen = o_gail.predict_reward(pd, rs, obs[0], actions[0], 0.99, masks[0], update_rms=False)
    np.testing.assert_allclose(frozen.numpy(), g["rewards_frozen"], rtol=1e-4, atol=1e-6)
