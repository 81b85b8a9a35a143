"""CPU: the Philox restatement of oracle/actor.py against the published known-answer vectors of Philox4x32-10
(Random123 kat_vectors), and the sampling oracle against the reference's distribution semantics."""
import numpy as np
import torch

from oracle import actor


def test_philox_known_answers():
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = actor.philox4x32_10(np.array([ctr], np.uint32), key)[0]
        assert tuple(int(x) for x in got) == want


def test_categorical_oracle_matches_torch_distribution():
    rng = np.random.default_rng(0)
    z = rng.normal(size=(4, 7)).astype(np.float32)
    z = np.repeat(z, 20000, 0)
    v, a, lp, u, cdf = actor.categorical(z, seed=1234567, offset=3)
    d = torch.distributions.Categorical(logits=torch.from_numpy(z[:, 1:]))
    np.testing.assert_allclose(lp, d.log_prob(torch.from_numpy(a)).numpy(), rtol=1e-5, atol=1e-6)
    p = d.probs.numpy()
    for r in range(4):
        sl = slice(r * 20000, (r + 1) * 20000)
        freq = np.bincount(a[sl], minlength=6) / 20000.0
        assert np.abs(freq - p[r * 20000]).max() < 0.015
    _, am, lpm, _, _ = actor.categorical(z[:8], 1, 0, deterministic=True)
    assert np.array_equal(am, z[:8, 1:].argmax(1))


def test_gaussian_oracle_matches_torch_distribution():
    rng = np.random.default_rng(1)
    z = np.repeat(rng.normal(size=(1, 4)).astype(np.float32), 50000, 0)
    logstd = np.array([-0.3, 0.1, 0.4], np.float32)
    v, a, lp = actor.gaussian(z, logstd, seed=99, offset=0)
    d = torch.distributions.Normal(torch.from_numpy(z[:, 1:]), torch.from_numpy(np.exp(logstd)))
    np.testing.assert_allclose(lp, d.log_prob(torch.from_numpy(a)).sum(-1).numpy(), rtol=1e-4, atol=1e-5)
    assert np.abs(a.mean(0) - z[0, 1:]).max() < 0.02
    assert np.abs(a.std(0) - np.exp(logstd)).max() < 0.02
    _, am, _ = actor.gaussian(z[:4], logstd, 1, 0, deterministic=True)
    assert np.array_equal(am, z[:4, 1:])
