"""Diagnostic for the peer-memory exchange on one GPU: where does a two-rank all-reduce stall?"""
import ctypes
import os
import sys
import time

os.environ["CUDA_DEVICE_MAX_CONNECTIONS"] = os.environ.get("A2CB_PROBE_CONN", "32")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-a2c-ppo-acktr-gail_b200"))
import torch
from cuda import cuda, cudart
from a2c_ppo_acktr import _dist, _native

DEV = "cuda:0"
torch.cuda.init()
torch.zeros(1, device=DEV)


def poll(streams, secs):
    t0 = time.time()
    while time.time() - t0 < secs:
        if all(s.query() for s in streams):
            return True
        time.sleep(0.01)
    return False


# ---- 1. does a stream memory wait work at all, and does it block only its own stream?  (A2CB_PROBE_MEMOP=1; on the
#         pool's driver the launch behind the pending wait blocks the HOST: the script never gets past it) ----
if os.environ.get("A2CB_PROBE_MEMOP") != "1":
    sys.argv.append("skip")
flag = torch.zeros(4, dtype=torch.int32, device=DEV)
s0, s1, s2 = torch.cuda.Stream(DEV), torch.cuda.Stream(DEV), torch.cuda.Stream(DEV)
torch.cuda.synchronize()
if "skip" not in sys.argv:
    out = torch.zeros(1, device=DEV)
    out.add_(1)
    flag.fill_(0)                              # both kernels loaded before anything waits (lazy module loading)
    torch.cuda.synchronize()
    r = cuda.cuStreamWaitValue32(s0.cuda_stream, flag.data_ptr(), 1, 0)
    print("cuStreamWaitValue32 rc", r, flush=True)
    with torch.cuda.stream(s0):
        out.add_(1)
    print("s0 done before flag (expect False):", poll([s0], 0.3), flush=True)
    with torch.cuda.stream(s1):
        flag.fill_(1)
    print("s1 done (expect True):", poll([s1], 2.0), " s0 done after flag (expect True):", poll([s0], 2.0), flush=True)

# ---- 2. two emulated ranks ----
world, count = 2, 1 << 16
mode = os.environ.get("A2CB_P2P_WAIT", "memop")
g = [_dist.PeerExchange(DEV, world, r, count) for r in range(world)]
_dist.PeerExchange.connect_local(g)
cap = (count + 4 * world + 255) // 256 * 256
bases = []
for x in g:
    b = ctypes.c_void_p()
    _native.lib().a2cb_p2p_local_base(x.handle, ctypes.byref(b))
    bases.append(b.value)
host = torch.zeros(2, 64, dtype=torch.int32).pin_memory()


def counters(tag):
    for r in range(world):
        cudart.cudaMemcpyAsync(host[r].data_ptr(), bases[r] + cap * 8, 256, cudart.cudaMemcpyKind.cudaMemcpyDeviceToHost, s2.cuda_stream)
    ok = poll([s2], 2.0)
    print(tag, "copy done", ok, "A/B/tickets rank0", host[0][[0, 16, 32, 33]].tolist(), "rank1", host[1][[0, 16, 32, 33]].tolist(), flush=True)


xs = [torch.full((count,), float(r + 1), device=DEV) for r in range(world)]
torch.cuda.synchronize()
counters("start (%s)" % mode)
g[0].all_reduce_sum(xs[0], stream=s0)
print("rank 0 enqueued; s0 done (expect False):", poll([s0], 0.5), flush=True)
counters("after rank 0")
g[1].all_reduce_sum(xs[1], stream=s1)
done = poll([s0, s1], 5.0)
print("both done:", done, flush=True)
counters("end")
if done:
    print("values", xs[0][:3].tolist(), xs[1][-3:].tolist(), flush=True)
    for rep in range(3):
        for r in range(world):
            g[r].all_reduce_sum(xs[r], stream=(s0, s1)[r])
        print("rep", rep, "done:", poll([s0, s1], 5.0), xs[0][0].item(), flush=True)
sys.stdout.flush()
os._exit(0)
