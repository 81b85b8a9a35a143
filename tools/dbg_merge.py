"""Per-tensor distance of one A2C / PPO update from the reference golden (A/B runs of an engine knob).

    A2CB_TCX_MERGE=0 python tools/dbg_merge.py C2 1
"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-a2c-ppo-acktr-gail_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import helpers
import test_update_gpu as T

name, u8 = sys.argv[1], bool(int(sys.argv[2]))
case, cfg, pol, st, agent = T.build(name, u8)
torch.manual_seed(7)
out = agent.update(st)
print("losses", out, "golden", case["losses"].tolist())
tot_n = tot_d = 0.0
for k, v in pol.named_parameters():
    want = case["finaldense_" + k].astype(np.float64)
    got = helpers.dense(v).astype(np.float64)
    n, d = float(((got - want) ** 2).sum()), float((want ** 2).sum())
    tot_n += n; tot_d += d
    print("%-28s rel %.2e  absdiff-rms %.2e  want-rms %.2e" % (k, (n / max(d, 1e-300)) ** 0.5, (n / got.size) ** 0.5, (d / got.size) ** 0.5))
print("all", (tot_n / tot_d) ** 0.5)
g = pol._flat_grad
print("grad norm", float(g.double().norm()), "grad checksum", float(g.double().sum()))
np.save(os.path.join("/tmp", "dbg_grad_%s_%s.npy" % (name, os.environ.get("A2CB_TCX_MERGE", "1"))), g.cpu().numpy())
rt = pol.runtime()
keys = sorted(rt.arena.bufs.keys())
if os.environ.get("DBG_KEYS") == "1":
    print(keys)
save = {}
for k in keys:
    base = k.split("/")[-1] if isinstance(k, str) else str(k)
    if any(base.endswith(s) for s in ("a1", "a1_lo", "a2", "a2_lo", "a3", "a3_lo", "h", "g1", "g2", "g3", "x0b")):
        save[base] = rt.arena.bufs[k].detach().float().cpu().numpy() if rt.arena.bufs[k].dtype != torch.uint8 else rt.arena.bufs[k].cpu().numpy()
np.savez(os.path.join("/tmp", "dbg_acts_%s_%s.npz" % (name, os.environ.get("A2CB_TCX_MERGE", "7"))), **save)
