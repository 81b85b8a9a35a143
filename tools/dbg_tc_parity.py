import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-a2c-ppo-acktr-gail_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import helpers
import test_update_gpu as T
from a2c_ppo_acktr import _native
for eng in (1, 0):
    _native.set_gemm_mode(eng)
    case, cfg, pol, st, agent = T.build("C3", True)
    first = {}
    def hook(step):
        if step == 1:
            first["grads"] = {k: p.grad.detach().clone() for k, p in pol.named_parameters()}
    agent.step_hook = hook
    torch.manual_seed(7)
    agent.update(st)
    print("engine", eng)
    for k, v in first["grads"].items():
        want = case["step1_graddense_" + k].astype(np.float64)
        got = helpers.dense(v).astype(np.float64)
        print("  %-28s relL2 %.3e  norm %.3e  maxabs %.3e" % (k, np.linalg.norm(got - want) / max(np.linalg.norm(want), 1e-30), np.linalg.norm(want), np.abs(got - want).max()))
