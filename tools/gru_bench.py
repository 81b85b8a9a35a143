"""Times the persistent GRU recurrence kernels (a2cb_gru_seq) per kernel family and checks them against each other.

    python tools/gru_bench.py [T N H [family]]

Prints one JSON line per family: ms per forward / backward launch (CUDA events, L2 flushed between launches) and the
rel-L2 distance of its outputs from the fp32 SIMT kernels.
"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-a2c-ppo-acktr-gail_b200"))

import torch  # noqa: E402

from a2c_ppo_acktr import _engine as E  # noqa: E402
from a2c_ppo_acktr import _native  # noqa: E402


def main():
    T, N, H = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (128, 64, 512)
    dev = "cuda:0"
    g = torch.Generator().manual_seed(5)
    gi = torch.randn(T * N, 3 * H, generator=g)
    Whh = torch.randn(3 * H, H, generator=g) / H ** 0.5
    bhh = torch.randn(3 * H, generator=g) * 0.1
    h0 = torch.randn(N, H, generator=g)
    masks = (torch.rand(T * N, generator=g) >= 0.01).float()
    dout = torch.randn(T * N, H, generator=g) * 1e-3
    dz = lambda *s: torch.zeros(*s, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    lib = _native.lib()
    ref = None
    only = sys.argv[4] if len(sys.argv) >= 5 else None            # e.g. "tc": time this family only (ncu captures)
    for mode, code in (("simt", 0), ("mma", 1), ("tc", 2)):
        if only and mode != only:
            continue
        _native.set_gru_mode(code)
        if _native.gru_seq_variant(N, H) != mode:
            print(json.dumps({"mode": mode, "skipped": "shape runs %s" % _native.gru_seq_variant(N, H)}))
            continue
        bufs = dict(gi=gi.to(dev), gh=dz(T * N, 3 * H), hm=dz(T * N, H), h0=h0.to(dev), masks=masks.to(dev),
                    out=dz(T * N, H), r=dz(T * N, H), z=dz(T * N, H), n=dz(T * N, H), dout=dout.to(dev),
                    dgi=dz(T * N, 3 * H), dgh=dz(T * N, 3 * H), dcarry=dz(N, H))
        W_d, b_d = Whh.to(dev), bhh.to(dev)
        d = E.GruSeqDesc()
        for k, v in bufs.items():
            setattr(d, k, v.data_ptr())
        d.T, d.N, d.H, d.t = T, N, H, 0
        d.W_hh, d.b_hh = W_d.data_ptr(), b_d.data_ptr()
        st = torch.cuda.current_stream().cuda_stream
        ms = {0: [], 1: []}
        for rep in range(6):
            for direction in (0, 1):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _native.check(lib.a2cb_gru_seq(direction, ctypes.addressof(d), st), "gru_seq")
                e1.record()
                torch.cuda.synchronize()
                if rep >= 2:
                    ms[direction].append(e0.elapsed_time(e1))
        outs = {k: bufs[k].clone() for k in ("out", "hm", "dgi", "dgh")}
        rec = {"mode": mode, "T": T, "N": N, "H": H, "fwd_ms": sum(ms[0]) / len(ms[0]), "bwd_ms": sum(ms[1]) / len(ms[1]),
               "fwd_us_per_step": 1e3 * sum(ms[0]) / len(ms[0]) / T, "bwd_us_per_step": 1e3 * sum(ms[1]) / len(ms[1]) / T}
        if ref is None:
            ref = outs
        else:
            for k in outs:
                rec["rel_" + k] = float((outs[k] - ref[k]).double().norm() / ref[k].double().norm().clamp_min(1e-30))
        print(json.dumps(rec), flush=True)
    _native.set_gru_mode(-1)
    if os.environ.get("A2CB_GRU_TC_DEBUG") == "1":
        import numpy as np
        buf = np.zeros((2, 16, 16), np.int64)
        _native.check(lib.a2cb_gru_tc_debug_stamps(ctypes.c_void_p(buf.ctypes.data)), "stamps")
        for direction, name, n in ((0, "fwd", 10), (1, "bwd", 9)):
            st = buf[direction, :, :n].astype(np.float64)
            prev_end = np.roll(st[:, n - 1], 1 if direction == 0 else -1)
            d = np.diff(st, axis=1)
            ok = slice(2, 14)
            print(json.dumps({"stamps": name, "phase_cycles_median": np.median(d[ok], axis=0).round(0).tolist(),
                              "step_cycles_median": float(np.median(np.abs(np.diff(st[:, 0]))[ok]))}))
            # extra stamps (slots n..15), relative to the step's first stamp: forward 11..13 = arrival of state chunks 1..3,
            # 10 = chunk 0's instructions issued
            ex = buf[direction, :, n:].astype(np.float64) - buf[direction, :, :1].astype(np.float64)
            print(json.dumps({"stamps": name + "_extra_since_stamp0", "slots": list(range(n, 16)),
                              "cycles_median": np.median(ex[ok], axis=0).round(0).tolist()}))
    for cs in (16, 8, 4, 2):
        print(json.dumps({"cluster_size": cs, "max_active_clusters_217KB": int(lib.a2cb_gru_cluster_probe(cs, 217 * 1024)),
                          "max_active_clusters_100KB": int(lib.a2cb_gru_cluster_probe(cs, 100 * 1024))}))


if __name__ == "__main__":
    main()
