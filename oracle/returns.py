"""Oracle: RolloutStorage.compute_returns restated as a sequential numpy recurrence.

Follows reference a2c_ppo_acktr/storage.py:66-105 operation by operation (each
numpy fp32 op rounds once, like the torch CPU kernels the reference dispatches;
the Python-double product gamma*gae_lambda is rounded to fp32 only when it meets
the array, as at storage.py:79).
"""
import numpy as np


def compute_returns(rewards, value_preds, masks, bad_masks, next_value, use_gae, gamma,
                    gae_lambda, use_proper_time_limits=True, dtype=np.float32):
    """Inputs are array-likes of shape [T,N,1] / [T+1,N,1] / [N,1].

    Returns (returns [T+1,N,1], value_preds [T+1,N,1]) -- the second because the
    GAE variants overwrite value_preds[T] (storage.py:74,93); the non-GAE variants
    write returns[T] (storage.py:85,102).  `dtype=np.float64` gives the
    high-precision tie-breaker run of the same recurrence.
    """
    r = np.asarray(rewards, dtype=dtype)
    v = np.array(value_preds, dtype=dtype, copy=True)
    m = np.asarray(masks, dtype=dtype)
    b = np.asarray(bad_masks, dtype=dtype)
    nv = np.asarray(next_value, dtype=dtype).reshape(v[-1].shape)
    T = r.shape[0]
    out = np.zeros_like(v)
    one = dtype(1.0)
    if use_gae:
        v[T] = nv
        gae = np.zeros_like(v[0])
        for t in range(T - 1, -1, -1):
            delta = r[t] + gamma * v[t + 1] * m[t + 1] - v[t]              # storage.py:76-78
            gae = delta + gamma * gae_lambda * m[t + 1] * gae               # storage.py:79-80
            if use_proper_time_limits:
                gae = gae * b[t + 1]                                        # storage.py:82
            out[t] = gae + v[t]                                             # storage.py:83
    else:
        out[T] = nv
        for t in range(T - 1, -1, -1):
            disc = out[t + 1] * gamma * m[t + 1] + r[t]                     # storage.py:104-105
            if use_proper_time_limits:
                disc = disc * b[t + 1] + (one - b[t + 1]) * v[t]            # storage.py:87-89
            out[t] = disc
    return out, v
