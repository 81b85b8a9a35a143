"""Generates tests/golden/*.npz by running the UNMODIFIED reference on the CPU.

Run in the authoring container only (needs /root/reference):
    python tests/golden/make_golden.py [case ...]
The reference is imported through oracle/ref_shim.py (stubbed gym/sb3 imports,
torch.symeig -> linalg.eigh); none of its code is copied.  Frames are not stored:
they are regenerated from the counter-based generator in synthetic.py, so every
fixture stays small.  torch 2.11.0 CPU, default thread count.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pytorch-a2c-ppo-acktr-gail_b200"))

import synthetic  # noqa: E402
from oracle import ref_shim  # noqa: E402

REF = ref_shim.install()
Discrete, Box = ref_shim.Discrete, ref_shim.Box

SAMPLE = 16


def summary(t):
    """Compact fingerprint of a tensor: fp64 sum, fp64 L2 norm, SAMPLE strided values."""
    f = t.detach().reshape(-1).double()
    n = f.numel()
    pos = torch.linspace(0, n - 1, min(SAMPLE, n)).long()
    return np.concatenate([[f.sum().item(), f.norm().item(), float(n)], f[pos].numpy()])


DENSE = 4096


def dense(t):
    """DENSE strided values of a tensor (fp32) -- enough for a rel-L2 estimate of large tensors."""
    f = t.detach().reshape(-1).float()
    n = f.numel()
    pos = torch.linspace(0, n - 1, min(DENSE, n)).long()
    return f[pos].numpy().copy()


def space(action):
    kind, n = action
    return Discrete(n) if kind == "discrete" else Box(n)


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrs.items()})
    print("wrote", path, os.path.getsize(path), "bytes")


# ---------------------------------------------------------------------------
def case_returns():
    out = {}
    shapes = [(5, 16), (20, 32), (128, 8), (2048, 1), (7, 3), (1, 1), (33, 130)]
    g = torch.Generator().manual_seed(11)
    for si, (T, N) in enumerate(shapes):
        rewards = torch.randn(T, N, 1, generator=g)
        values = torch.randn(T + 1, N, 1, generator=g)
        masks = (torch.rand(T + 1, N, 1, generator=g) > 0.1).float()
        bad = (torch.rand(T + 1, N, 1, generator=g) > 0.05).float()
        if si == 0:
            masks.fill_(1.0)            # no episode boundary at all
        if si == 4:
            masks.fill_(0.0)            # every step terminal
        nv = torch.randn(N, 1, generator=g)
        out["in%d_rewards" % si] = rewards.numpy()
        out["in%d_values" % si] = values.numpy()
        out["in%d_masks" % si] = masks.numpy()
        out["in%d_bad" % si] = bad.numpy()
        out["in%d_next" % si] = nv.numpy()
        for use_gae in (0, 1):
            for ptl in (0, 1):
                st = REF["storage"].RolloutStorage(T, N, (3,), Discrete(2), 1)
                st.rewards.copy_(rewards)
                st.value_preds.copy_(values)
                st.masks.copy_(masks)
                st.bad_masks.copy_(bad)
                st.compute_returns(nv, bool(use_gae), 0.99, 0.95, bool(ptl))
                out["out%d_g%d_p%d_returns" % (si, use_gae, ptl)] = st.returns.numpy().copy()
                out["out%d_g%d_p%d_vlast" % (si, use_gae, ptl)] = st.value_preds[-1].numpy().copy()
    out["shapes"] = np.array(shapes)
    save("returns", **out)


def _fill_storage(st, g):
    for name in ("obs", "recurrent_hidden_states", "rewards", "value_preds", "returns", "action_log_probs",
                 "masks"):
        t = getattr(st, name)
        t.copy_(torch.randn(t.shape, generator=g))
    if st.actions.dtype == torch.int64:
        st.actions.copy_(torch.randint(0, 6, st.actions.shape, generator=g))
    else:
        st.actions.copy_(torch.randn(st.actions.shape, generator=g))


def case_generators():
    out = {}
    g = torch.Generator().manual_seed(5)
    names = ["obs", "hxs", "actions", "values", "returns", "masks", "oldlp", "adv"]
    # feed-forward, discrete actions, num_mini_batch and explicit mini_batch_size (drop_last)
    for tag, (T, N, kw, sp) in {
        "ffA": (8, 4, dict(num_mini_batch=4), Discrete(6)),
        "ffB": (7, 3, dict(mini_batch_size=5), Box(2)),
    }.items():
        st = REF["storage"].RolloutStorage(T, N, (2, 3, 3), sp, 5)
        _fill_storage(st, g)
        adv = torch.randn(T, N, 1, generator=g)
        for name in ("obs", "recurrent_hidden_states", "rewards", "value_preds", "returns",
                     "action_log_probs", "actions", "masks"):
            out["%s_st_%s" % (tag, name)] = getattr(st, name).numpy().copy()
        out[tag + "_adv"] = adv.numpy()
        torch.manual_seed(7)
        batches = list(st.feed_forward_generator(adv, **kw))
        out[tag + "_nb"] = np.array(len(batches))
        for b, tup in enumerate(batches):
            for name, x in zip(names, tup):
                out["%s_b%d_%s" % (tag, b, name)] = x.numpy().copy()
        # advantages=None path used by GAIL (reference gail.py:60-61)
        torch.manual_seed(7)
        tup = next(iter(st.feed_forward_generator(None, **kw)))
        assert tup[-1] is None
    # recurrent
    T, N = 6, 8
    st = REF["storage"].RolloutStorage(T, N, (2, 3, 3), Discrete(6), 5)
    _fill_storage(st, g)
    adv = torch.randn(T, N, 1, generator=g)
    for name in ("obs", "recurrent_hidden_states", "rewards", "value_preds", "returns",
                 "action_log_probs", "actions", "masks"):
        out["rec_st_%s" % name] = getattr(st, name).numpy().copy()
    out["rec_adv"] = adv.numpy()
    torch.manual_seed(7)
    batches = list(st.recurrent_generator(adv, 4))
    out["rec_nb"] = np.array(len(batches))
    for b, tup in enumerate(batches):
        for name, x in zip(names, tup):
            out["rec_b%d_%s" % (b, name)] = x.numpy().copy()
    save("generators", **out)


POLICY_CASES = {
    "cnn": dict(obs_shape=(4, 84, 84), action=("discrete", 6), recurrent=False, B=4),
    "mlp": dict(obs_shape=(11,), action=("box", 3), recurrent=False, B=6),
    "cnn_gru": dict(obs_shape=(4, 84, 84), action=("discrete", 6), recurrent=True, B=12, N=3),
    "mlp_gru": dict(obs_shape=(11,), action=("box", 3), recurrent=True, B=12, N=3),
    "mlp_disc": dict(obs_shape=(5,), action=("discrete", 4), recurrent=False, B=6),
}


def policy_inputs(pc, seed=3):
    B = pc["B"]
    if len(pc["obs_shape"]) == 3:
        obs = torch.from_numpy(synthetic.atari_frames(1, B, seed, None, pc["obs_shape"])[0].astype(np.float32))
    else:
        obs = torch.from_numpy(synthetic.hopper_obs(1, B, seed, None, pc["obs_shape"][0])[0])
    g = torch.Generator().manual_seed(seed)
    kind, na = pc["action"]
    action = torch.randint(0, na, (B, 1), generator=g) if kind == "discrete" else torch.randn(B, na, generator=g)
    masks = (torch.rand(B, 1, generator=g) > 0.3).float()
    return obs, action, masks, g


def case_policy():
    out = {}
    for tag, pc in POLICY_CASES.items():
        torch.manual_seed(1)
        pol = REF["model"].Policy(pc["obs_shape"], space(pc["action"]), base_kwargs={"recurrent": pc["recurrent"]})
        for k, v in pol.state_dict().items():
            out["%s_init_%s" % (tag, k)] = summary(v)
        obs, action, masks, g = policy_inputs(pc)
        H = pol.recurrent_hidden_state_size
        nh = pc.get("N", pc["B"])
        hxs = torch.randn(nh, H, generator=g) if pc["recurrent"] else torch.zeros(nh, H)
        v, lp, ent, h2 = pol.evaluate_actions(obs, hxs, masks, action)
        out[tag + "_value"] = v.detach().numpy()
        out[tag + "_logp"] = lp.detach().numpy()
        out[tag + "_entropy"] = ent.detach().numpy()
        out[tag + "_hxs_out"] = h2.detach().numpy()
        out[tag + "_hxs_in"] = hxs.numpy()
        # deterministic act on single-step input
        hx1 = torch.randn(pc["B"], H, generator=g) if pc["recurrent"] else torch.zeros(pc["B"], H)
        with torch.no_grad():
            v1, a1, lp1, h1 = pol.act(obs, hx1, masks, deterministic=True)
            torch.manual_seed(9)
            v2, a2, lp2, _ = pol.act(obs, hx1, masks, deterministic=False)
        out[tag + "_act_hxs_in"] = hx1.numpy()
        out[tag + "_act_value"] = v1.numpy()
        out[tag + "_act_action"] = a1.numpy()
        out[tag + "_act_logp"] = lp1.numpy()
        out[tag + "_act_hxs_out"] = h1.numpy()
        out[tag + "_sample_action"] = a2.numpy()
        out[tag + "_sample_logp"] = lp2.numpy()
    save("policy", **out)


UPDATE_CASES = {
    # name: (base config, overrides)
    "C1s": ("C1", dict(T=64, N=1, ppo_epoch=2, num_mini_batch=4)),
    "C3s": ("C3", dict(T=8, N=4, ppo_epoch=2, num_mini_batch=2)),
    "C3u": ("C3", dict(T=8, N=4, ppo_epoch=1, num_mini_batch=2, use_clipped_value_loss=False,
                       use_proper_time_limits=True)),
    "C5s": ("C5", dict(T=8, N=4, ppo_epoch=2, num_mini_batch=2)),
    # the benchmarked shape: one GPU's C5 shard, T = 128, 256 environments, recurrent minibatches of 64 environments
    # (8,192 rows); one epoch keeps the reference run short.  Only slots 0 and T of the hidden states are stored (the
    # update reads slot 0 only, reference storage.py:160-161; after_update() copies slot T into it).
    "C5m": ("C5", dict(T=128, N=256, ppo_epoch=1, num_mini_batch=4)),
    "C2": ("C2", dict()),
    "C2g": ("C2", dict(T=6, N=3, use_gae=True, use_proper_time_limits=True)),
    "C1": ("C1", dict()),
    "C3": ("C3", dict()),
    "C4s": ("C4", dict(T=4, N=4)),
    "C4": ("C4", dict()),
    # ACKTR on the MLP policy (MLPBase + DiagGaussian; 13 KFAC modules incl. the logstd AddBias)
    "C4m": ("C1", dict(algo="acktr", T=32, N=4, entropy_coef=0.01, use_gae=False, use_proper_time_limits=False)),
}


def plain_name(k):
    """KFACOptimizer rewrites the module tree (SplitBias, kfac.py:101-108): map its parameter names
    back to the original ones so that ACKTR fixtures use the same keys as every other case."""
    return k.replace(".module.weight", ".weight").replace(".add_bias._bias", ".bias")


def plain_value(k, v):
    return v.reshape(-1) if k.endswith(".add_bias._bias") else v



def build_ref_rollout(cfg, pol, seed=0):
    H = pol.recurrent_hidden_state_size
    ro = synthetic.make_rollout(
        cfg, lambda o, h, m: pol.act(o, h, m), lambda o, h, m: pol.get_value(o, h, m), H, seed)
    return ro


def storage_from(ro, cfg, H):
    st = REF["storage"].RolloutStorage(cfg["T"], cfg["N"], cfg["obs_shape"], space(cfg["action"]), H)
    for name in ("obs", "recurrent_hidden_states", "rewards", "value_preds", "action_log_probs", "actions",
                 "masks", "bad_masks"):
        getattr(st, name).copy_(ro[name])
    return st


def case_update(name):
    base, over = UPDATE_CASES[name]
    cfg = dict(synthetic.CONFIGS[base])
    cfg.update(over)
    torch.manual_seed(1)
    pol = REF["model"].Policy(cfg["obs_shape"], space(cfg["action"]), base_kwargs={"recurrent": cfg["recurrent"]})
    H = pol.recurrent_hidden_state_size
    torch.manual_seed(2)                      # action sampling stream
    ro = build_ref_rollout(cfg, pol, seed=0)
    st = storage_from(ro, cfg, H)
    st.compute_returns(ro["next_value"], cfg["use_gae"], cfg["gamma"], cfg["gae_lambda"],
                       cfg["use_proper_time_limits"])
    out = {"config": np.array(json.dumps(cfg))}
    for k in ("actions", "action_log_probs", "value_preds", "next_value"):
        out["ro_" + k] = ro[k].numpy()
    if cfg["recurrent"]:
        if cfg["T"] * cfg["N"] > 4096:
            out["ro_h0"] = ro["recurrent_hidden_states"][0].numpy()
            out["ro_hT"] = ro["recurrent_hidden_states"][-1].numpy()      # after_update() carries it into slot 0
        else:
            out["ro_recurrent_hidden_states"] = ro["recurrent_hidden_states"].numpy()
    out["returns"] = st.returns.numpy().copy()
    out["obs_sum"] = np.array(ro["obs"].double().sum().item())
    for k, v in pol.state_dict().items():
        out["init_" + k] = summary(v)

    if cfg["algo"] == "acktr":
        agent = REF["algo.a2c_acktr"].A2C_ACKTR(pol, cfg["value_loss_coef"], cfg["entropy_coef"], acktr=True)
    elif cfg["algo"] == "ppo":
        agent = REF["algo.ppo"].PPO(pol, cfg["clip_param"], cfg["ppo_epoch"], cfg["num_mini_batch"],
                                    cfg["value_loss_coef"], cfg["entropy_coef"], lr=cfg["lr"], eps=cfg["eps"],
                                    max_grad_norm=cfg["max_grad_norm"],
                                    use_clipped_value_loss=cfg.get("use_clipped_value_loss", True))
    else:
        agent = REF["algo.a2c_acktr"].A2C_ACKTR(pol, cfg["value_loss_coef"], cfg["entropy_coef"], lr=cfg["lr"],
                                                eps=cfg["eps"], alpha=cfg["alpha"],
                                                max_grad_norm=cfg["max_grad_norm"])
    # first-step fingerprints: hook the optimiser to capture clipped grads and params after step 1
    first = {}
    orig_step = agent.optimizer.step

    def step_hook(*a, **k):
        if "grads" not in first:
            first["grads"] = {plain_name(n): plain_value(n, p.grad.detach().clone()) for n, p in pol.named_parameters()}
        r = orig_step(*a, **k)
        if "params" not in first:
            first["params"] = {plain_name(n): plain_value(n, p.detach().clone()) for n, p in pol.named_parameters()}
        return r

    agent.optimizer.step = step_hook
    torch.manual_seed(7)
    vl, al, ent = agent.update(st)
    out["losses"] = np.array([vl, al, ent], dtype=np.float64)
    if cfg["algo"] == "acktr":
        # running Kronecker factors after the first update, keyed by the parameter each module owns
        names = {id(p): plain_name(n) for n, p in pol.named_parameters()}
        for mod in agent.optimizer.modules:
            key = names[id(next(mod.parameters()))]
            out["maa_" + key] = summary(agent.optimizer.m_aa[mod])
            out["mgg_" + key] = summary(agent.optimizer.m_gg[mod])
            out["maadense_" + key] = dense(agent.optimizer.m_aa[mod])
            out["mggdense_" + key] = dense(agent.optimizer.m_gg[mod])
    for k, v in first["grads"].items():
        out["step1_grad_" + k] = summary(v)
        out["step1_graddense_" + k] = dense(v)
    for k, v in first["params"].items():
        out["step1_param_" + k] = summary(v)
        out["step1_paramdense_" + k] = dense(v)
    small = sum(p.numel() for p in pol.parameters()) < 20000
    for k0, v0 in pol.state_dict().items():
        k, v = plain_name(k0), plain_value(k0, v0)
        out["final_" + k] = summary(v)
        out["finaldense_" + k] = dense(v)
        if small:
            out["finalfull_" + k] = v.numpy().copy()
    # second update on the same storage (optimizer state carried over)
    st.after_update()
    torch.manual_seed(8)
    vl2, al2, ent2 = agent.update(st)
    out["losses2"] = np.array([vl2, al2, ent2], dtype=np.float64)
    save("update_" + name, **out)


def gail_inputs(seed=21, T=32, N=4, obs_dim=11, act_dim=3, n_expert=80):
    """Seeded inputs of the GAIL case (shared with tests/test_gail_*.py through the fixture itself)."""
    g = torch.Generator().manual_seed(seed)
    return dict(obs=torch.randn(T + 1, N, obs_dim, generator=g), actions=torch.randn(T, N, act_dim, generator=g),
                masks=(torch.rand(T + 1, N, 1, generator=g) > 0.1).float(),
                expert_states=torch.randn(n_expert, obs_dim, generator=g) * 2.0 + 0.3,
                expert_actions=torch.randn(n_expert, act_dim, generator=g))


def gail_obsfilt(x, update=False):
    """Stand-in for VecNormalize._obfilt (reference envs.py:199-208 with frozen statistics): numpy in, numpy out."""
    return np.clip((x - 0.3) / np.sqrt(4.0 + 1e-8), -10.0, 10.0)


def case_gail():
    """reference algo/gail.py: two Discriminator.update calls, then the per-step predict_reward loop of main.py:152-155."""
    gail = REF["algo.gail"]
    inp = gail_inputs()
    T, N = inp["actions"].shape[:2]
    obs_dim, act_dim = inp["obs"].shape[-1], inp["actions"].shape[-1]
    torch.manual_seed(1)
    disc = gail.Discriminator(obs_dim + act_dim, 100, torch.device("cpu"))
    out = {"in_" + k: v.numpy() for k, v in inp.items()}
    for k, v in disc.state_dict().items():
        out["init_" + k] = v.numpy().copy()
    st = REF["storage"].RolloutStorage(T, N, (obs_dim,), Box(act_dim), 1)
    st.obs.copy_(inp["obs"])
    st.actions.copy_(inp["actions"])
    st.masks.copy_(inp["masks"])
    ds = torch.utils.data.TensorDataset(inp["expert_states"], inp["expert_actions"])
    loader = torch.utils.data.DataLoader(ds, batch_size=16, shuffle=True, drop_last=True)
    torch.manual_seed(7)
    losses = [disc.update(loader, st, gail_obsfilt) for _ in range(2)]
    out["losses"] = np.array(losses, np.float64)
    for k, v in disc.state_dict().items():
        out["final_" + k] = v.numpy().copy()
    rewards = torch.zeros(T, N, 1)
    for step in range(T):
        rewards[step] = disc.predict_reward(st.obs[step], st.actions[step], 0.99, st.masks[step])
    out["rewards"] = rewards.numpy()
    out["rms"] = np.array([float(np.asarray(disc.ret_rms.mean).reshape(-1)[0]), float(np.asarray(disc.ret_rms.var).reshape(-1)[0]),
                           float(disc.ret_rms.count)], np.float64)
    out["returns_state"] = disc.returns.numpy().copy()
    # a second rollout's worth with update_rms=False (evaluation path)
    out["rewards_frozen"] = disc.predict_reward(st.obs[0], st.actions[0], 0.99, st.masks[0], update_rms=False).numpy()
    save("gail", **out)


if __name__ == "__main__":
    want = sys.argv[1:] or ["returns", "generators", "policy", "gail"] + ["update:" + k for k in UPDATE_CASES]
    for w in want:
        if w == "returns":
            case_returns()
        elif w == "generators":
            case_generators()
        elif w == "policy":
            case_policy()
        elif w == "gail":
            case_gail()
        elif w.startswith("update:"):
            case_update(w.split(":", 1)[1])
        else:
            raise SystemExit("unknown case " + w)
