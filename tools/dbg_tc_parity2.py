import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-a2c-ppo-acktr-gail_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import helpers
import test_update_gpu as T
from a2c_ppo_acktr import _native, _engine
from a2c_ppo_acktr.algo import ppo as ppo_mod
_native.set_gemm_mode(0)
case, cfg, pol, st, agent = T.build("C3", True)
rt = pol.runtime()
adv = ppo_mod._adv_normalize(rt, st)
accum = rt.arena.ensure("loss_accum", 4)
prog, _ = agent._program(rt, st, 256, adv, accum)
torch.manual_seed(7)
idx = torch.randperm(1024)[:256].to("cuda:0")
# keep only fwd/loss/bwd ops (drop norm/optim)
ops = [(k, f, d) for (k, f, d) in prog.entries if k in (_engine.OP_GEMM, _engine.OP_REDUCE, _engine.OP_LOSS)]
names = []
gi = 0
for k, f, d in ops:
    if k == _engine.OP_GEMM:
        names.append(prog.plan_names[gi]); gi += 1
    else:
        names.append("op%d" % k)
def run(tc_index):
    for i, (k, f, d) in enumerate(ops):
        if k == _engine.OP_GEMM:
            d.tile = 1000 if i == tc_index else 0
        p = _engine.Program(torch.device("cuda:0")); p.add(k, d, f); p.run(idx)
    torch.cuda.synchronize()
    return pol._flat_grad.clone()
base = run(-1)
L = pol._layout
for i, n in enumerate(names):
    if ops[i][0] != _engine.OP_GEMM:
        continue
    g = run(i)
    out = []
    for lay in ("conv1", "conv2", "conv3", "fc", "heads"):
        w, b = L.layers[lay]
        d = (g[w:b] - base[w:b]).double().norm() / base[w:b].double().norm()
        out.append("%s %.1e" % (lay, d.item()))
    print("%-14s -> %s" % (n, "  ".join(out)))
