"""Relative error of both GEMM engines against fp64 as a function of the reduction length K."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-a2c-ppo-acktr-gail_b200"))
from a2c_ppo_acktr import _engine, _native
from a2c_ppo_acktr import _plans as P
dev = torch.device("cuda:0")
for kind in ("normal", "positive"):
    for K in (32, 128, 512, 2048, 8192, 32768):
        M, N = 256, 64
        g = torch.Generator().manual_seed(K)
        A = torch.randn(M, K, generator=g); B = torch.randn(K, N, generator=g)
        if kind == "positive":
            A, B = A.abs(), B.abs()
        ref = (A.double() @ B.double())
        out = {}
        for eng in (1, 0):
            _native.set_gemm_mode(eng)
            arena = _engine.Arena(dev)
            arena.bind("A", A.reshape(-1).to(dev)); arena.bind("B", B.reshape(-1).to(dev)); arena.bind("C", torch.zeros(M * N, device=dev))
            plan = P.Plan("t", ("A", 0), ("B", 0), ("C", 0), M, N, K, P.plain(K), P.plain(1), P.plain(N), P.plain(1), P.plain(N), P.plain(1), vecA=1, vecB=1)
            prog = _engine.Program(dev); prog.add_plan(plan, arena); prog.run(); torch.cuda.synchronize()
            C = arena.bufs["C"].cpu().double().view(M, N)
            out[eng] = ((C - ref).norm() / ref.norm()).item(), ((C - ref).abs().max() / ref.abs().max()).item(), ((C-ref).mean()/ref.abs().mean()).item()
        t32 = (A @ B).double()
        print("%-8s K=%6d  tc: relL2 %.2e max %.2e bias %+.2e | simt: relL2 %.2e max %.2e | torch-cpu fp32 relL2 %.2e" % (
            kind, K, out[1][0], out[1][1], out[1][2], out[0][0], out[0][1], ((t32 - ref).norm() / ref.norm()).item()))
