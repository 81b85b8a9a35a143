"""Kernel microbenchmarks (CUDA events, L2 flushed between iterations): scan and gather GB/s."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-a2c-ppo-acktr-gail_b200"))
from a2c_ppo_acktr import _native  # noqa: E402

dev = torch.device("cuda:0")
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def scan(T, N, gae, ptl, schedule):
    r = torch.randn(T, N, 1, device=dev)
    v = torch.randn(T + 1, N, 1, device=dev)
    m = (torch.rand(T + 1, N, 1, device=dev) > 0.01).float()
    b = torch.ones(T + 1, N, 1, device=dev)
    out = torch.zeros(T + 1, N, 1, device=dev)
    nv = torch.randn(N, device=dev)
    med, best = timeit(lambda: _native.returns_scan(r, v, out, m, b, nv, 0.99, 0.95, gae, ptl, schedule))
    bpt = {(1, 1): 20, (0, 1): 20, (1, 0): 16, (0, 0): 12}[(gae, ptl)]
    gbs = bpt * T * N / med / 1e6
    print(json.dumps(dict(kernel="returns_scan", T=T, N=N, gae=gae, ptl=ptl, schedule=schedule, ms=med, ms_best=best,
                          GBs=gbs, frac=gbs / PEAK)))


def gather(rows, row_bytes, n_src):
    src = torch.randint(0, 255, (n_src, row_bytes), dtype=torch.uint8, device=dev)
    dst = torch.empty(rows, row_bytes, dtype=torch.uint8, device=dev)
    idx = torch.randint(0, n_src, (rows,), device=dev)
    med, best = timeit(lambda: _native.gather_rows([(src, dst)], idx, rows))
    gbs = 2 * rows * row_bytes / med / 1e6
    print(json.dumps(dict(kernel="gather_rows", rows=rows, row_bytes=row_bytes, ms=med, ms_best=best, GBs=gbs,
                          frac=gbs / PEAK)))


if __name__ == "__main__":
    for (T, N) in [(128, 1 << 20), (128, 2048), (128, 8), (2048, 1)]:
        for gae, ptl in [(1, 1), (1, 0), (0, 0)]:
            for schedule in ([1] if N >= (1 << 20) else [1, 2]):
                scan(T, N, gae, ptl, schedule)
    gather(256, 28224, 1024)
    gather(65536, 28224, 262144 // 2)
    gather(16384, 112896, 32768)
