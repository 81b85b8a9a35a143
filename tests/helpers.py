"""Shared helpers of the test-suite: golden loading, rollout reconstruction, fingerprints."""
import json
import os

import numpy as np
import torch

import synthetic

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SAMPLE = 16


def load(name):
    with np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def summary(t):
    f = t.detach().cpu().reshape(-1).double()
    n = f.numel()
    pos = torch.linspace(0, n - 1, min(SAMPLE, n)).long()
    return np.concatenate([[f.sum().item(), f.norm().item(), float(n)], f[pos].numpy()])


DENSE = 4096


def dense(t):
    f = t.detach().cpu().reshape(-1).float()
    n = f.numel()
    pos = torch.linspace(0, n - 1, min(DENSE, n)).long()
    return f[pos].numpy().copy()


def rel_l2(case, prefix, named):
    """rel-L2 distance ||got - want|| / ||want|| over the dense samples of ALL tensors together
    (the parameter-vector metric of SURVEY.md section 8c)."""
    num = den = 0.0
    for k, v in named.items():
        want = case[prefix + k].astype(np.float64)
        got = dense(v).astype(np.float64)
        num += float(((got - want) ** 2).sum())
        den += float((want ** 2).sum())
    return (num / max(den, 1e-300)) ** 0.5


def case_config(case):
    cfg = json.loads(str(case["config"]))
    cfg["obs_shape"] = tuple(cfg["obs_shape"])
    cfg["action"] = tuple(cfg["action"])
    return cfg


def hidden_size_of(cfg):
    if not cfg["recurrent"]:
        return 1
    return 512 if len(cfg["obs_shape"]) == 3 else 64


_INPUT_CACHE = {}        # the counter-based generator costs ~1 min for a C5-sized rollout: reuse across parametrisations


def rollout_from_case(case, cfg, obs_uint8=False, device="cpu"):
    """Rebuilds the rollout the golden file was produced from: frames/signals from the
    counter-based generator, policy outputs (actions, log-probs, values) from the file."""
    key = (cfg["kind"], cfg["T"], cfg["N"], tuple(cfg["obs_shape"]), bool(obs_uint8))
    if key not in _INPUT_CACHE:
        if len(_INPUT_CACHE) >= 2:
            _INPUT_CACHE.clear()
        _INPUT_CACHE[key] = synthetic.rollout_inputs(cfg, seed=0, obs_uint8=obs_uint8)
    obs, rewards, masks, bad = (a.copy() if a.nbytes < (64 << 20) else a for a in _INPUT_CACHE[key])
    T, N = cfg["T"], cfg["N"]
    H = hidden_size_of(cfg)
    ro = {
        "obs": torch.from_numpy(obs),
        "rewards": torch.from_numpy(rewards),
        "masks": torch.from_numpy(masks),
        "bad_masks": torch.from_numpy(bad),
        "actions": torch.from_numpy(case["ro_actions"]),
        "action_log_probs": torch.from_numpy(case["ro_action_log_probs"]),
        "value_preds": torch.from_numpy(case["ro_value_preds"]).clone(),
        "next_value": torch.from_numpy(case["ro_next_value"]),
        "returns": torch.zeros(T + 1, N, 1),
    }
    if cfg["recurrent"] and "ro_h0" in case:
        # large recurrent cases store slot 0 only: the update reads nothing else (reference storage.py:160-161)
        ro["recurrent_hidden_states"] = torch.zeros(T + 1, N, H)
        ro["recurrent_hidden_states"][0].copy_(torch.from_numpy(case["ro_h0"]))
        if "ro_hT" in case:
            ro["recurrent_hidden_states"][-1].copy_(torch.from_numpy(case["ro_hT"]))
    elif cfg["recurrent"]:
        ro["recurrent_hidden_states"] = torch.from_numpy(case["ro_recurrent_hidden_states"])
    else:
        ro["recurrent_hidden_states"] = torch.zeros(T + 1, N, H)
    assert abs(ro["obs"].double().sum().item() - float(case["obs_sum"])) < 1e-3, "synthetic frames drifted"
    return {k: v.to(device) for k, v in ro.items()}


def assert_summary_close(got, want, rtol, atol, what="", elem_frac=0.0):
    """Compares two `summary` fingerprints (sum, norm, n, samples).
    elem_frac widens the per-sample tolerance by that fraction of the tensor's RMS value."""
    got, want = np.asarray(got), np.asarray(want)
    assert got[2] == want[2], "%s: element count %s != %s" % (what, got[2], want[2])
    n = want[2]
    # the L2 norm bounds a sum's rounding noise: |sum err| <= sqrt(n) * |elem err|
    scale = max(want[1], 1e-30)
    assert abs(got[1] - want[1]) <= rtol * scale + atol, "%s: norm %r vs %r" % (what, got[1], want[1])
    assert abs(got[0] - want[0]) <= ((rtol + elem_frac) * scale + atol) * np.sqrt(n), "%s: sum %r vs %r" % (what, got[0], want[0])
    np.testing.assert_allclose(got[3:], want[3:], rtol=rtol, atol=atol + (rtol + elem_frac) * scale / np.sqrt(n), err_msg=what)


def relu_knife_edge(case, cfg, obs_uint8, rel=3e-6):
    """Smallest |pre-activation| / layer RMS of the CNN trunk's LAST ReLU layer (Linear 1568 -> 512, reference
    model.py:176-188) on the case's update batch, computed with the fp32 CPU oracle parameters, and whether it is
    below `rel`.

    A unit whose pre-activation is that close to zero relative to the layer's scale has an UNDETERMINED sign at fp32
    accuracy: the reference's own rounding, the exact-fp32 engine and the 3xTF32 tensor-core engine may each put it on
    either side, and its relu' bit then switches a whole row-gradient entry of the 512-wide feature layer on or off
    (every layer below inherits it).  The small goldens sit at >= 1.7e-4 (80 x 512 units: nothing near zero) -- except
    update_C2, whose unit [72, 188] has 8.0e-7 against an RMS of 0.55 (1.5e-6); the full-size C3 case has 1.1e-6.  Update
    goldens with such a unit cannot be held to the 1e-4 vector bound on the parameters after the step (losses are)."""
    import torch.nn.functional as F
    from oracle import policy as o_pol
    if len(cfg["obs_shape"]) != 3:
        return 1.0, False
    ro = rollout_from_case(case, cfg, obs_uint8=obs_uint8)
    spec = o_pol.Spec(cfg["obs_shape"], cfg["action"], cfg["recurrent"])
    p = o_pol.init_params(spec, seed=1)
    a = ro["obs"][:-1].reshape(-1, *cfg["obs_shape"]).float() / 255.0
    with torch.no_grad():
        for i, stride in ((0, 4), (2, 2), (4, 1)):
            a = F.conv2d(a, p["base.main.%d.weight" % i], p["base.main.%d.bias" % i], stride=stride).relu()
        z = F.linear(a.reshape(a.shape[0], -1), p["base.main.7.weight"], p["base.main.7.bias"])
        worst = float(z.abs().min() / (z ** 2).mean().sqrt())
    return worst, worst < rel
