"""Per-layer timing of one PPO minibatch step on both GEMM engines.
usage: python tools/layer_bench.py [B ...]   (A2CB_ENGINES=tc,simt)"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-a2c-ppo-acktr-gail_b200"))
from a2c_ppo_acktr import _engine, _native, algo  # noqa: E402
from a2c_ppo_acktr.model import Policy  # noqa: E402
from a2c_ppo_acktr.storage import RolloutStorage  # noqa: E402


class Discrete(object):
    n = 6


def main():
    dev = torch.device("cuda:0")
    Bs = [int(x) for x in sys.argv[1:]] or [256, 8192]
    engines = os.environ.get("A2CB_ENGINES", "tc,simt").split(",")
    reps = int(os.environ.get("REPS", "5"))
    for B in Bs:
        T, N = 128, max(8, B * 4 // 128)
        torch.manual_seed(1)
        pol = Policy((4, 84, 84), Discrete()).to(dev)
        st = RolloutStorage(T, N, (4, 84, 84), Discrete(), 1, obs_dtype=torch.uint8)
        st.to(dev)
        st.obs.random_(0, 256)
        st.actions.random_(0, 6)
        st.returns.normal_()
        st.value_preds.normal_()
        st.action_log_probs.fill_(-1.79)
        agent = algo.PPO(pol, 0.1, 4, 4, 0.5, 0.01, lr=2.5e-4, eps=1e-5, max_grad_norm=0.5)
        rt = pol.runtime()
        adv = rt.arena.ensure("adv", T * N)
        adv.normal_()
        accum = rt.arena.ensure("loss_accum", 4)
        prog, _ = agent._program(rt, st, B, adv, accum)
        singles = []
        for kind, flags, desc in prog.entries:
            p = _engine.Program(dev)
            p.add(kind, desc, flags)
            singles.append(p.finalize())
        labels, gi = [], 0
        for kind, flags, desc in prog.entries:
            if kind == _engine.OP_GEMM:
                labels.append(prog.plan_names[gi]); gi += 1
            else:
                labels.append({2: "reduce_splits", 3: "loss", 4: "grad_norm", 5: "optim", 10: "bias_psum", 11: "weight_prep"}.get(kind, "op%d" % kind))
        idx = torch.randperm(T * N)[:B].to(dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        for eng in engines:
            _native.set_gemm_mode(1 if eng == "tc" else 0)
            times = np.zeros(len(singles))
            for r in range(reps + 2):
                agent.optimizer.configure(prog.optim_desc)
                flush.zero_()
                evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(singles) + 1)]
                evs[0].record()
                for i, p in enumerate(singles):
                    p.run(idx)
                    evs[i + 1].record()
                torch.cuda.synchronize()
                if r >= 2:
                    times += np.array([evs[i].elapsed_time(evs[i + 1]) for i in range(len(singles))])
            times /= reps
            tot = times.sum()
            print("== B=%d engine=%s total %.3f ms (%.1f TFLOP/s algorithmic)" % (B, eng, tot, prog.flops / tot / 1e9))
            for (kind, flags, desc), lab, t in zip(prog.entries, labels, times):
                if kind == _engine.OP_GEMM:
                    f = 2.0 * desc.M * desc.N * desc.K * desc.batch
                    print("  %-14s M=%-8d N=%-4d K=%-8d z=%d*%-3d %8.3f ms %7.2f TFLOP/s" % (
                        lab, desc.M, desc.N, desc.K, desc.batch, desc.splits, t, f / t / 1e9))
                else:
                    print("  %-14s %47s %8.3f ms" % (lab, "", t))


if __name__ == "__main__":
    main()
