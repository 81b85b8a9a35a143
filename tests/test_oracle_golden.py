"""Pins the CPU oracle (oracle/) against the golden files produced by the unmodified
reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

import helpers
from oracle import losses, minibatch, policy, returns, update

SHAPES = 7


@pytest.mark.parametrize("si", range(SHAPES))
@pytest.mark.parametrize("use_gae", [0, 1])
@pytest.mark.parametrize("ptl", [0, 1])
def test_returns_bit_exact(si, use_gae, ptl):
    g = helpers.load("returns")
    out, v = returns.compute_returns(g["in%d_rewards" % si], g["in%d_values" % si], g["in%d_masks" % si],
                                     g["in%d_bad" % si], g["in%d_next" % si], bool(use_gae), 0.99, 0.95, bool(ptl))
    want = g["out%d_g%d_p%d_returns" % (si, use_gae, ptl)]
    T = want.shape[0] - 1
    # returns[T] is only defined (written) by the non-GAE variants
    upto = T if use_gae else T + 1
    assert np.array_equal(out[:upto], want[:upto])
    assert np.array_equal(v[-1], g["out%d_g%d_p%d_vlast" % (si, use_gae, ptl)])


def _storage_dict(g, tag):
    names = ("obs", "recurrent_hidden_states", "rewards", "value_preds", "returns", "action_log_probs",
             "actions", "masks")
    return {n: torch.from_numpy(g["%s_st_%s" % (tag, n)]) for n in names}


NAMES = ["obs", "hxs", "actions", "values", "returns", "masks", "oldlp", "adv"]


@pytest.mark.parametrize("tag,kw", [("ffA", dict(num_mini_batch=4)), ("ffB", dict(mini_batch_size=5))])
def test_feed_forward_generator_bit_exact(tag, kw):
    g = helpers.load("generators")
    ro = _storage_dict(g, tag)
    adv = torch.from_numpy(g[tag + "_adv"])
    T, N = ro["rewards"].shape[:2]
    torch.manual_seed(7)
    chunks = minibatch.feed_forward_index_chunks(T, N, **kw)
    assert len(chunks) == int(g[tag + "_nb"])
    for b, idx in enumerate(chunks):
        tup = minibatch.gather_feed_forward(ro, idx, adv)
        for name, x in zip(NAMES, tup):
            assert np.array_equal(x.numpy(), g["%s_b%d_%s" % (tag, b, name)]), (tag, b, name)


def test_recurrent_generator_bit_exact():
    g = helpers.load("generators")
    ro = _storage_dict(g, "rec")
    adv = torch.from_numpy(g["rec_adv"])
    torch.manual_seed(7)
    chunks = minibatch.recurrent_env_chunks(ro["rewards"].shape[1], 4)
    assert len(chunks) == int(g["rec_nb"])
    for b, env in enumerate(chunks):
        tup = minibatch.gather_recurrent(ro, env, adv)
        for name, x in zip(NAMES, tup):
            assert np.array_equal(x.numpy(), g["rec_b%d_%s" % (b, name)]), (b, name)


def test_recurrent_generator_ragged_quirk():
    # 10 envs / 4 minibatches -> 5 chunks of 2; 9 envs / 4 -> IndexError (storage.py:154,165)
    torch.manual_seed(0)
    assert len(minibatch.recurrent_env_chunks(10, 4)) == 5
    with pytest.raises(IndexError):
        minibatch.recurrent_env_chunks(9, 4)


POLICY_CASES = {
    "cnn": dict(obs_shape=(4, 84, 84), action=("discrete", 6), recurrent=False, B=4),
    "mlp": dict(obs_shape=(11,), action=("box", 3), recurrent=False, B=6),
    "cnn_gru": dict(obs_shape=(4, 84, 84), action=("discrete", 6), recurrent=True, B=12, N=3),
    "mlp_gru": dict(obs_shape=(11,), action=("box", 3), recurrent=True, B=12, N=3),
    "mlp_disc": dict(obs_shape=(5,), action=("discrete", 4), recurrent=False, B=6),
}


def policy_inputs(pc, seed=3):
    import synthetic
    B = pc["B"]
    if len(pc["obs_shape"]) == 3:
        obs = torch.from_numpy(synthetic.atari_frames(1, B, seed, None, pc["obs_shape"])[0].astype(np.float32))
    else:
        obs = torch.from_numpy(synthetic.hopper_obs(1, B, seed, None, pc["obs_shape"][0])[0])
    gen = torch.Generator().manual_seed(seed)
    kind, na = pc["action"]
    action = torch.randint(0, na, (B, 1), generator=gen) if kind == "discrete" else torch.randn(B, na, generator=gen)
    masks = (torch.rand(B, 1, generator=gen) > 0.3).float()
    return obs, action, masks, gen


@pytest.mark.parametrize("tag", sorted(POLICY_CASES))
def test_policy_matches_reference(tag):
    g = helpers.load("policy")
    pc = POLICY_CASES[tag]
    spec = policy.Spec(pc["obs_shape"], pc["action"], pc["recurrent"])
    p = policy.init_params(spec, seed=1)
    for k, v in p.items():                      # identical initialisation stream
        np.testing.assert_allclose(helpers.summary(v), g["%s_init_%s" % (tag, k)], rtol=1e-5, atol=1e-7, err_msg=k)
    obs, action, masks, gen = policy_inputs(pc)
    hxs = torch.from_numpy(g[tag + "_hxs_in"])
    with torch.no_grad():
        v, lp, ent, h2 = policy.evaluate_actions(p, spec, obs, hxs, masks, action)
    tol = dict(rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(v.numpy(), g[tag + "_value"], **tol)
    np.testing.assert_allclose(lp.numpy(), g[tag + "_logp"], **tol)
    np.testing.assert_allclose(ent.numpy(), g[tag + "_entropy"], **tol)
    np.testing.assert_allclose(h2.numpy(), g[tag + "_hxs_out"], **tol)
    hx1 = torch.from_numpy(g[tag + "_act_hxs_in"])
    with torch.no_grad():
        v1, a1, lp1, h1 = policy.act(p, spec, obs, hx1, masks, deterministic=True)
        torch.manual_seed(9)
        _, a2, lp2, _ = policy.act(p, spec, obs, hx1, masks, deterministic=False)
    np.testing.assert_allclose(v1.numpy(), g[tag + "_act_value"], **tol)
    np.testing.assert_allclose(a1.numpy(), g[tag + "_act_action"], **tol)
    np.testing.assert_allclose(lp1.numpy(), g[tag + "_act_logp"], **tol)
    np.testing.assert_allclose(h1.numpy(), g[tag + "_act_hxs_out"], **tol)
    np.testing.assert_allclose(a2.numpy(), g[tag + "_sample_action"], **tol)     # same RNG consumption
    np.testing.assert_allclose(lp2.numpy(), g[tag + "_sample_logp"], **tol)


UPDATE_SMALL = ["C1s", "C3s", "C3u", "C5s", "C2", "C2g", "C5m"]      # C5m: the benchmarked shape (128 x 256, 4 steps)
UPDATE_FULL = ["C1", "C3"]


def run_update_case(name):
    case = helpers.load("update_" + name)
    cfg = helpers.case_config(case)
    spec = policy.Spec(cfg["obs_shape"], cfg["action"], cfg["recurrent"])
    p0 = policy.init_params(spec, seed=1)
    for k, v in p0.items():
        np.testing.assert_allclose(helpers.summary(v), case["init_" + k], rtol=1e-5, atol=1e-7, err_msg=k)
    ro = helpers.rollout_from_case(case, cfg)
    ret, vp = returns.compute_returns(ro["rewards"].numpy(), ro["value_preds"].numpy(), ro["masks"].numpy(),
                                      ro["bad_masks"].numpy(), ro["next_value"].numpy(), cfg["use_gae"],
                                      cfg["gamma"], cfg["gae_lambda"], cfg["use_proper_time_limits"])
    T = cfg["T"]
    upto = T if cfg["use_gae"] else T + 1
    assert np.array_equal(ret[:upto], case["returns"][:upto])
    ro["returns"] = torch.from_numpy(ret)
    ro["value_preds"] = torch.from_numpy(vp)
    p = update.make_leaf_params(p0)
    trace = []
    torch.manual_seed(7)
    if cfg["algo"] == "ppo":
        vl, al, ent, opt = update.ppo_update(p, spec, ro, cfg, trace=trace)
    else:
        vl, al, ent, opt = update.a2c_update(p, spec, ro, cfg, trace=trace)
    return case, cfg, p, (vl, al, ent), trace


@pytest.mark.parametrize("name", UPDATE_SMALL + UPDATE_FULL)
def test_update_matches_reference(name):
    case, cfg, p, got, trace = run_update_case(name)
    want = case["losses"]
    # the oracle runs the same torch kernels in the same order as the reference: agreement is
    # at rounding level on the first step and stays far below the 1e-4 bar for these sizes
    np.testing.assert_allclose(np.array(got), want, rtol=2e-4, atol=2e-6)
    for k, v in trace[0]["grads"].items():
        helpers.assert_summary_close(helpers.summary(v), case["step1_grad_" + k], rtol=1e-4, atol=1e-7, what="grad " + k)
    # Full-size updates (16 / 320 optimiser steps) are chaotic: the reference differs from ITSELF
    # by 1e-4..3e-4 rel-L2 on the whole parameter vector when only the thread count changes
    # (SURVEY.md H1), more on the near-zero bias tensors.  Small cases (<= 4 steps) stay tight.
    # Individual Adam-updated bias entries move by +-lr per step on sign flips of tiny gradients,
    # so single samples of near-zero tensors get a 10%-of-RMS allowance there.
    full = name in UPDATE_FULL
    for k, v in p.items():
        helpers.assert_summary_close(helpers.summary(v), case["final_" + k], rtol=1e-2 if full else 2e-4,
                                     atol=1e-6, what="param " + k, elem_frac=0.1 if full else 0.0)


@pytest.mark.parametrize("name", ["C4s", "C4", "C4m"])
def test_acktr_oracle_matches_reference(name):
    """oracle/kfac.py (functional KFAC) against the reference's hook-based KFACOptimizer."""
    from oracle import kfac
    case = helpers.load("update_" + name)
    cfg = helpers.case_config(case)
    spec = policy.Spec(cfg["obs_shape"], cfg["action"], cfg["recurrent"])
    p0 = policy.init_params(spec, seed=1)
    ro = helpers.rollout_from_case(case, cfg)
    ret, vp = returns.compute_returns(ro["rewards"].numpy(), ro["value_preds"].numpy(), ro["masks"].numpy(),
                                      ro["bad_masks"].numpy(), ro["next_value"].numpy(), cfg["use_gae"],
                                      cfg["gamma"], cfg["gae_lambda"], cfg["use_proper_time_limits"])
    ro["returns"], ro["value_preds"] = torch.from_numpy(ret), torch.from_numpy(vp)
    p = update.make_leaf_params(p0)
    st = kfac.KFACState()
    torch.manual_seed(7)
    got = kfac.acktr_update(p, ro, cfg, st)
    np.testing.assert_allclose(np.array(got), case["losses"], rtol=1e-4, atol=2e-6)
    for k in p:
        for kind, m in (("maa", st.m_aa), ("mgg", st.m_gg)):
            want = case["%sdense_%s" % (kind, k)].astype(np.float64)
            g = helpers.dense(m[k]).astype(np.float64)
            assert np.linalg.norm(g - want) <= 1e-4 * max(np.linalg.norm(want), 1e-30), (kind, k)
    assert helpers.rel_l2(case, "finaldense_", p) <= 1e-4
    for k in ("obs", "recurrent_hidden_states", "masks", "bad_masks"):
        ro[k][0].copy_(ro[k][-1])
    torch.manual_seed(8)
    got2 = kfac.acktr_update(p, ro, cfg, st)
    np.testing.assert_allclose(np.array(got2), case["losses2"], rtol=5e-4, atol=5e-6)
