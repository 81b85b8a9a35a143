"""Deterministic synthetic rollouts of the BASELINE.json config shapes.

Counter-based (splitmix64) so that any slice of the data -- e.g. one GPU's
environment shard -- can be produced independently and bit-identically on any
machine, without shipping frames as fixtures.  Everything here is plain numpy on
the host; neither the oracle nor the CUDA path is involved.

Shapes follow SURVEY.md section 8d:
  Atari-shaped (C2-C5): obs uint8 uniform 0..255 [T+1, N, 4, 84, 84];
      masks = 1 - Bernoulli(0.01); bad_masks = 1; rewards in {-1, 0, +1}, P(+-1) = 0.02 each.
  Hopper-shaped (C1):   obs ~ N(0,1) clipped to +-10 [T+1, N, 11]; rewards ~ N(1, 0.5);
      masks = 1 - Bernoulli(0.002); bad_masks = 1 - Bernoulli(0.001).
Actions / log-probs / value predictions are NOT produced here: a policy-consistent
rollout needs `Policy.act`, which the caller runs (reference main.py:113-134).
"""
import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def hash_u64(index, stream):
    """64-bit hash of (index array, stream id)."""
    with np.errstate(over="ignore"):
        idx = np.asarray(index, dtype=np.uint64)
        return _splitmix(_splitmix(idx) ^ _splitmix(np.uint64(stream) * np.uint64(0xD1342543DE82EF95) + np.uint64(1)))


def uniform01(index, stream):
    return (hash_u64(index, stream) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def normal(index, stream):
    u1 = uniform01(index, stream * 2 + 1)
    u2 = uniform01(index, stream * 2 + 2)
    return np.sqrt(-2.0 * np.log(1.0 - u1)) * np.cos(2.0 * np.pi * u2)


def _grid(T, n_lo, n_hi, N_total, inner):
    """Global element counters for rows t in [0,T), envs [n_lo,n_hi) of an [T, N_total, inner] tensor."""
    t = np.arange(T, dtype=np.uint64)[:, None, None]
    n = np.arange(n_lo, n_hi, dtype=np.uint64)[None, :, None]
    k = np.arange(inner, dtype=np.uint64)[None, None, :]
    return (t * np.uint64(N_total) + n) * np.uint64(inner) + k


def atari_frames(T1, N_total, seed=0, env_range=None, obs_shape=(4, 84, 84), chunk_t=8):
    """uint8 [T1, n, *obs_shape] for envs in env_range (default all)."""
    lo, hi = env_range if env_range is not None else (0, N_total)
    inner = int(np.prod(obs_shape))
    out = np.empty((T1, hi - lo, inner), dtype=np.uint8)
    for t0 in range(0, T1, chunk_t):
        t1 = min(T1, t0 + chunk_t)
        t = np.arange(t0, t1, dtype=np.uint64)[:, None, None]
        n = np.arange(lo, hi, dtype=np.uint64)[None, :, None]
        k = np.arange(inner, dtype=np.uint64)[None, None, :]
        ctr = (t * np.uint64(N_total) + n) * np.uint64(inner) + k
        out[t0:t1] = (hash_u64(ctr, seed * 16 + 0) >> np.uint64(56)).astype(np.uint8)
    return out.reshape((T1, hi - lo) + tuple(obs_shape))


def atari_signals(T, N_total, seed=0, env_range=None):
    """rewards [T,n,1], masks [T+1,n,1], bad_masks [T+1,n,1] as fp32."""
    lo, hi = env_range if env_range is not None else (0, N_total)
    u = uniform01(_grid(T, lo, hi, N_total, 1), seed * 16 + 1)
    rewards = np.where(u < 0.02, 1.0, np.where(u < 0.04, -1.0, 0.0)).astype(np.float32)
    um = uniform01(_grid(T + 1, lo, hi, N_total, 1), seed * 16 + 2)
    masks = (um >= 0.01).astype(np.float32)
    bad = np.ones_like(masks)
    return rewards, masks, bad


def hopper_obs(T1, N_total, seed=0, env_range=None, obs_dim=11):
    lo, hi = env_range if env_range is not None else (0, N_total)
    z = normal(_grid(T1, lo, hi, N_total, obs_dim), seed * 16 + 3)
    return np.clip(z, -10.0, 10.0).astype(np.float32)


def hopper_signals(T, N_total, seed=0, env_range=None):
    lo, hi = env_range if env_range is not None else (0, N_total)
    rewards = (1.0 + 0.5 * normal(_grid(T, lo, hi, N_total, 1), seed * 16 + 4)).astype(np.float32)
    masks = (uniform01(_grid(T + 1, lo, hi, N_total, 1), seed * 16 + 5) >= 0.002).astype(np.float32)
    bad = (uniform01(_grid(T + 1, lo, hi, N_total, 1), seed * 16 + 6) >= 0.001).astype(np.float32)
    return rewards, masks, bad


# The five BASELINE.json configurations (SURVEY.md section 8 shorthand C1..C5).
CONFIGS = {
    "C1": dict(kind="hopper", algo="ppo", obs_shape=(11,), action=("box", 3), T=2048, N=1,
               recurrent=False, ppo_epoch=10, num_mini_batch=32, clip_param=0.2, lr=3e-4, eps=1e-5,
               value_loss_coef=0.5, entropy_coef=0.0, max_grad_norm=0.5, gamma=0.99, gae_lambda=0.95,
               use_gae=True, use_proper_time_limits=True),
    "C2": dict(kind="atari", algo="a2c", obs_shape=(4, 84, 84), action=("discrete", 6), T=5, N=16,
               recurrent=False, lr=7e-4, eps=1e-5, alpha=0.99, value_loss_coef=0.5, entropy_coef=0.01,
               max_grad_norm=0.5, gamma=0.99, gae_lambda=0.95, use_gae=False, use_proper_time_limits=False),
    "C3": dict(kind="atari", algo="ppo", obs_shape=(4, 84, 84), action=("discrete", 6), T=128, N=8,
               recurrent=False, ppo_epoch=4, num_mini_batch=4, clip_param=0.1, lr=2.5e-4, eps=1e-5,
               value_loss_coef=0.5, entropy_coef=0.01, max_grad_norm=0.5, gamma=0.99, gae_lambda=0.95,
               use_gae=True, use_proper_time_limits=False),
    "C4": dict(kind="atari", algo="acktr", obs_shape=(4, 84, 84), action=("discrete", 6), T=20, N=32,
               recurrent=False, value_loss_coef=0.5, entropy_coef=0.01, gamma=0.99, gae_lambda=0.95,
               use_gae=False, use_proper_time_limits=False),
    "C5": dict(kind="atari", algo="ppo", obs_shape=(4, 84, 84), action=("discrete", 6), T=128, N=2048,
               recurrent=True, ppo_epoch=4, num_mini_batch=4, clip_param=0.1, lr=2.5e-4, eps=1e-5,
               value_loss_coef=0.5, entropy_coef=0.01, max_grad_norm=0.5, gamma=0.99, gae_lambda=0.95,
               use_gae=True, use_proper_time_limits=False),
}


def rollout_inputs(cfg, seed=0, env_range=None, N_total=None, obs_uint8=False):
    """Host arrays (obs, rewards, masks, bad_masks) of one configuration (or env shard)."""
    T = cfg["T"]
    N_total = N_total if N_total is not None else cfg["N"]
    if cfg["kind"] == "atari":
        obs = atari_frames(T + 1, N_total, seed, env_range, cfg["obs_shape"])
        if not obs_uint8:
            obs = obs.astype(np.float32)
        rewards, masks, bad = atari_signals(T, N_total, seed, env_range)
    else:
        obs = hopper_obs(T + 1, N_total, seed, env_range, cfg["obs_shape"][0])
        rewards, masks, bad = hopper_signals(T, N_total, seed, env_range)
    return obs, rewards, masks, bad


def make_rollout(cfg, act_fn, value_fn, hidden_state_size, seed=0, env_range=None, N_total=None,
                 obs_uint8=False, device="cpu", inputs=None):
    """Policy-consistent rollout as a dict named like the RolloutStorage attributes.

    act_fn(obs, hxs, masks) -> (value, action, log_prob, hxs); value_fn(obs, hxs, masks) -> value.
    Mirrors the actor loop of reference main.py:113-139 on synthetic observations.
    """
    import torch
    # `inputs`: (obs, rewards, masks, bad_masks) already generated by the caller (same generator, same arguments)
    obs, rewards, masks, bad = inputs if inputs is not None else rollout_inputs(cfg, seed, env_range, N_total, obs_uint8)
    T, n = cfg["T"], obs.shape[1]
    ro = {
        "obs": torch.from_numpy(obs).to(device),
        "rewards": torch.from_numpy(rewards).to(device),
        "masks": torch.from_numpy(masks).to(device),
        "bad_masks": torch.from_numpy(bad).to(device),
        "recurrent_hidden_states": torch.zeros(T + 1, n, hidden_state_size, device=device),
        "value_preds": torch.zeros(T + 1, n, 1, device=device),
        "returns": torch.zeros(T + 1, n, 1, device=device),
        "action_log_probs": torch.zeros(T, n, 1, device=device),
    }
    kind, na = cfg["action"]
    ro["actions"] = (torch.zeros(T, n, 1, dtype=torch.int64, device=device) if kind == "discrete"
                     else torch.zeros(T, n, na, device=device))
    with torch.no_grad():
        for t in range(T):
            v, a, lp, h = act_fn(ro["obs"][t], ro["recurrent_hidden_states"][t], ro["masks"][t])
            ro["value_preds"][t].copy_(v)
            ro["actions"][t].copy_(a)
            ro["action_log_probs"][t].copy_(lp)
            ro["recurrent_hidden_states"][t + 1].copy_(h)
        ro["next_value"] = value_fn(ro["obs"][T], ro["recurrent_hidden_states"][T], ro["masks"][T]).detach()
    return ro
