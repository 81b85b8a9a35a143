import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-a2c-ppo-acktr-gail_b200"))
from a2c_ppo_acktr import _engine, _native
from a2c_ppo_acktr import _plans as P
dev = torch.device("cuda:0")
for (M, N, K, vA, vB, scale) in [(256, 64, 288, 1, 0, 1.0), (256, 64, 288, 1, 0, 1e-5), (256, 64, 256, 1, 0, 1e-5), (256, 64, 288, 1, 2, 1e-5),
                                 (256, 32, 256, 1, 0, 1e-6), (256, 64, 512, 1, 0, 1e-3), (256, 64, 288, 0, 0, 1e-5)]:
    g = torch.Generator().manual_seed(K)
    A = torch.randn(M, K, generator=g) * scale; B = torch.randn(N, K, generator=g)   # B stored [N][K]
    A = A * (torch.rand(M, K, generator=g) > 0.5)
    ref = (A.double() @ B.double().t())
    res = {}
    for eng in (1, 0):
        _native.set_gemm_mode(eng)
        arena = _engine.Arena(dev)
        arena.bind("A", A.reshape(-1).to(dev)); arena.bind("B", B.reshape(-1).to(dev)); arena.bind("C", torch.zeros(M * N, device=dev))
        plan = P.Plan("t", ("A", 0), ("B", 0), ("C", 0), M, N, K, P.plain(K), P.plain(1), P.plain(1), P.plain(K), P.plain(N), P.plain(1), vecA=vA, vecB=vB)
        prog = _engine.Program(dev); prog.add_plan(plan, arena); prog.run(); torch.cuda.synchronize()
        C = arena.bufs["C"].cpu().double().view(M, N)
        res[eng] = ((C - ref).norm() / ref.norm()).item()
    print("M=%d N=%d K=%d vecA=%d vecB=%d scale=%g  tc relL2 %.2e  simt relL2 %.2e" % (M, N, K, vA, vB, scale, res[1], res[0]))
