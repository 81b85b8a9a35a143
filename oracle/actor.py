"""Oracle: the sampling head of `Policy.act` (numpy).  TEST INFRASTRUCTURE ONLY.

Reference semantics: model.py:54-66 (`dist.mode()` when deterministic, else `dist.sample()`; then
`dist.log_probs(action)`), distributions.py:18-32 (FixedCategorical: multinomial sample, argmax mode, log-softmax
gathered at the action), :36-44 (FixedNormal: mean + std * N(0,1), mode = mean, summed log-density).
The reference draws from torch's device generator; the product draws from its own counter-based stream
(Philox4x32-10, Salmon et al. SC'11 -- the published algorithm restated here), so parity of the SAMPLED path is
pinned two ways: exactly against this restatement of the same stream, and statistically against the reference's
distribution (tests/test_actor_gpu.py).  The deterministic path is pinned to the reference goldens
(tests/golden/policy_*.npz, `act_*` entries).
"""
import numpy as np

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(counter, key):
    """counter: [..., 4] uint32, key: (k0, k1) -> [..., 4] uint32 (10 rounds)."""
    c = [counter[..., i].astype(np.uint64) for i in range(4)]
    k0, k1 = int(key[0]), int(key[1])
    for _ in range(10):
        p0, p1 = np.uint64(M0) * c[0], np.uint64(M1) * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & np.uint64(MASK)
        hi1, lo1 = p1 >> np.uint64(32), p1 & np.uint64(MASK)
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return np.stack(c, -1).astype(np.uint32)


def u01(x):
    """24 random bits -> the open interval (0, 1), as fp32."""
    return ((x >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)


def _counters(B, draw, offset, row_base=0):
    c = np.zeros((B, 4), np.uint32)
    c[:, 0] = (np.arange(B, dtype=np.uint64) + np.uint64(row_base)).astype(np.uint32)
    c[:, 1] = draw
    c[:, 2], c[:, 3] = offset & MASK, (offset >> 32) & MASK
    return c


def categorical(z, seed, offset, deterministic=False, row_base=0):
    """z: [B, 1 + A] fp32 head outputs.  -> (value [B], action [B] int64, logp [B], u [B], cdf [B, A])."""
    z = np.asarray(z, np.float32)
    B, A = z.shape[0], z.shape[1] - 1
    logits = z[:, 1:]
    mx = logits.max(1, keepdims=True)
    lse = mx[:, 0] + np.log(np.exp(logits - mx).sum(1, dtype=np.float32))
    p = np.exp(logits - lse[:, None]).astype(np.float32)
    cdf = np.cumsum(p, 1, dtype=np.float32)
    u = u01(philox4x32_10(_counters(B, 0, offset, row_base), (seed & MASK, (seed >> 32) & MASK))[:, 0])
    if deterministic:
        act = logits.argmax(1)
    else:
        act = np.minimum((u[:, None] >= cdf).sum(1), A - 1)          # first j with u < cdf_j, else the last one
    logp = logits[np.arange(B), act] - lse
    return z[:, 0], act.astype(np.int64), logp, u, cdf


def gaussian(z, logstd, seed, offset, deterministic=False, row_base=0):
    """z: [B, 1 + A] (value, mean), logstd [A].  -> (value, action [B, A], logp [B])."""
    z, logstd = np.asarray(z, np.float32), np.asarray(logstd, np.float32)
    B, A = z.shape[0], z.shape[1] - 1
    mean, std = z[:, 1:], np.exp(logstd)
    n = np.zeros((B, (A + 3) // 4 * 4), np.float32)
    if not deterministic:
        for blk in range((A + 3) // 4):
            r = philox4x32_10(_counters(B, blk, offset, row_base), (seed & MASK, (seed >> 32) & MASK))
            u = u01(r)
            r0, r1 = np.sqrt(-2.0 * np.log(u[:, 0])), np.sqrt(-2.0 * np.log(u[:, 2]))
            a0, a1 = 2.0 * np.pi * u[:, 1].astype(np.float64), 2.0 * np.pi * u[:, 3].astype(np.float64)
            n[:, 4 * blk + 0], n[:, 4 * blk + 1] = r0 * np.cos(a0), r0 * np.sin(a0)
            n[:, 4 * blk + 2], n[:, 4 * blk + 3] = r1 * np.cos(a1), r1 * np.sin(a1)
    act = (mean + std * n[:, :A]).astype(np.float32)
    diff = act - mean
    logp = (-(diff * diff) / (2.0 * std * std) - np.log(std) - np.float32(0.918938533204672742)).sum(1)
    return z[:, 0], act, logp.astype(np.float32)
