"""Frame re-layout kernel (tcx_s2d_atari_kernel) alone: ms per 8,192-image chunk for the A2CB_S2D_BY of this process,
plus checksums of both outputs (must not depend on the setting).  python tools/s2d_bench.py"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-a2c-ppo-acktr-gail_b200"))
import torch
from a2c_ppo_acktr import _engine, _native  # noqa: F401

DEV = "cuda:0"
B = 8192
g = torch.Generator().manual_seed(3)
frames = torch.randint(0, 256, (B, 4, 84, 84), generator=g, dtype=torch.uint8).to(DEV)
out = torch.empty(B, 21, 21, 64, device=DEV)
out_b = torch.empty(B, 21, 21, 64, device=DEV, dtype=torch.bfloat16)
lib = _native.lib()
st = _native.stream_ptr(DEV)


def run():
    _native.check(lib.a2cb_tcx_s2d_b16(out.data_ptr(), None, out_b.data_ptr(), frames.data_ptr(), 1, None, B, 4, 84, 84, 4, st), "s2d")


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
want = frames[:64].float().view(64, 4, 21, 4, 21, 4).permute(0, 2, 4, 1, 3, 5).reshape(64, 21, 21, 64)
print(json.dumps({"by_per_cta": os.environ.get("A2CB_S2D_BY", "default"), "ms": ms,
                  "write_TBps": (out.numel() * 4 + out_b.numel() * 2) / ms / 1e9,
                  "exact": bool(torch.equal(out[:64], want) and torch.equal(out_b[:64].float(), want)),
                  "sum": float(out.double().sum()), "sum_b16": float(out_b.double().sum())}))
