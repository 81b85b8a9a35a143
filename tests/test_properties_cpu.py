"""Property tests (hypothesis) of the host logic and the oracle that need no GPU: the migration plan of the balanced
sharded update against its loop definition on random shapes, and domain invariants of the returns recurrence
(reference storage.py:66-105) that the GPU parity tests rely on at full size."""
import numpy as np
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

from a2c_ppo_acktr import _dist
from oracle import returns as o_ret
from test_dist_cpu import _plan_loop


@settings(max_examples=60, deadline=None)
@given(world=st.sampled_from([2, 3, 4, 8]), per_rank_target=st.integers(1, 6), nmb=st.integers(1, 4),
       epochs=st.integers(1, 4), seed=st.integers(0, 10_000))
def test_balance_update_plan_matches_its_definition_on_random_shapes(world, per_rank_target, nmb, epochs, seed):
    n_local = per_rank_target * nmb                         # every rank gets `per_rank_target` environments per minibatch
    n_total = n_local * world
    g = torch.Generator().manual_seed(seed)
    perms = [torch.randperm(n_total, generator=g).numpy() for _ in range(epochs)]
    E_g = n_total // nmb
    got = _dist.balance_update_plan(perms, E_g, nmb, world, n_local)
    want = _plan_loop(perms, E_g, nmb, world, n_local)
    for r in range(world):
        assert got[r]["n_guests"] == want[r]["n_guests"]
        assert got[r]["first_send"] == want[r]["first_send"] and got[r]["first_recv"] == want[r]["first_recv"]
        for d in range(world):
            assert list(got[r]["send"][d]) == want[r]["send"][d]
            assert list(got[r]["recv"][d]) == want[r]["recv"][d]
        for e in range(epochs):
            for k in range(nmb):
                assert list(got[r]["envs"][e][k]) == want[r]["envs"][e][k]
    # conservation: what is sent is received, nobody sends to itself, a rank never both sends and receives in a minibatch
    for r in range(world):
        assert len(got[r]["send"][r]) == 0 and len(got[r]["recv"][r]) == 0
        assert sum(len(x) for x in got[r]["recv"]) == got[r]["n_guests"]
    assert sum(len(x) for p in got for x in p["send"]) == sum(p["n_guests"] for p in got)


def _rollout(T, N, seed, done_p):
    g = np.random.default_rng(seed)
    r = g.standard_normal((T, N, 1)).astype(np.float32)
    v = g.standard_normal((T + 1, N, 1)).astype(np.float32)
    m = (g.random((T + 1, N, 1)) > done_p).astype(np.float32)
    b = (g.random((T + 1, N, 1)) > done_p / 2).astype(np.float32)
    nv = g.standard_normal((N, 1)).astype(np.float32)
    return r, v, m, b, nv


@settings(max_examples=40, deadline=None)
@given(T=st.integers(1, 24), N=st.integers(1, 5), seed=st.integers(0, 10_000), use_gae=st.booleans(), ptl=st.booleans(),
       done_p=st.sampled_from([0.0, 0.1, 0.5]))
def test_returns_are_independent_per_environment_and_restart_at_episode_ends(T, N, seed, use_gae, ptl, done_p):
    """(i) environments do not interact: the returns of one column equal the run on that column alone (what lets the
    scan shard over environments without a collective); (ii) with an episode end between t and t + 1 (mask 0) nothing
    after t + 1 reaches returns[t] (storage.py:76-80, 104-105)."""
    r, v, m, b, nv = _rollout(T, N, seed, done_p)
    full, _ = o_ret.compute_returns(r, v, m, b, nv, use_gae, 0.99, 0.95, ptl)
    n = seed % N
    alone, _ = o_ret.compute_returns(r[:, n:n + 1], v[:, n:n + 1], m[:, n:n + 1], b[:, n:n + 1], nv[n:n + 1], use_gae, 0.99, 0.95, ptl)
    assert np.array_equal(full[:, n:n + 1], alone)
    cut = seed % T                                            # force an episode end after step `cut` of column n
    m2 = m.copy()
    m2[cut + 1, n] = 0.0
    base, _ = o_ret.compute_returns(r, v, m2, b, nv, use_gae, 0.99, 0.95, ptl)
    r3, v3, nv3 = r.copy(), v.copy(), nv.copy()
    r3[cut + 1:, n] += 7.0
    v3[cut + 2:, n] -= 3.0
    nv3[n] += 11.0
    if use_gae:
        # delta_t keeps the TERM -v[t] of its own step only; v[cut + 1] enters through gamma * v * mask = 0
        v3[cut + 1, n] -= 3.0
    moved, _ = o_ret.compute_returns(r3, v3, m2, b, nv3, use_gae, 0.99, 0.95, ptl)
    assert np.array_equal(base[:cut + 1, n], moved[:cut + 1, n])


@settings(max_examples=30, deadline=None)
@given(T=st.integers(1, 16), N=st.integers(1, 4), seed=st.integers(0, 10_000), ptl=st.booleans())
def test_gae_with_lambda_one_and_zero_values_is_the_discounted_return(T, N, seed, ptl):
    """GAE(gamma, lambda = 1) on zero value predictions telescopes to the discounted reward sum -- the non-GAE recurrence
    fed the same rollout with a zero bootstrap (float64 runs of both recurrences, no time limits)."""
    r, v, m, b, nv = _rollout(T, N, seed, 0.2)
    z, ones = np.zeros_like(v), np.ones_like(b)
    gae, _ = o_ret.compute_returns(r, z, m, ones, np.zeros_like(nv), True, 0.97, 1.0, ptl, dtype=np.float64)
    disc, _ = o_ret.compute_returns(r, z, m, ones, np.zeros_like(nv), False, 0.97, 1.0, ptl, dtype=np.float64)
    np.testing.assert_allclose(gae[:T], disc[:T], rtol=1e-12, atol=1e-12)
