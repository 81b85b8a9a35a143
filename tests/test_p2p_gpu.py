"""Peer-memory all-reduce (csrc/p2p.cu, a2cb_p2p_*) on ONE GPU: `world` exchange contexts live in this process, wired
to each other with a2cb_p2p_set_peers, every emulated rank on its own stream -- the hand-shakes (a counter bumped by the
peers' last CTA, awaited by a stream memory operation or one polling warp) are the ones the real multi-process runs use.  The cross-process
mapping (CUDA IPC) is covered by tests/dist_check.py under torchrun on 2+ GPUs."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from a2c_ppo_acktr import _dist, _native  # noqa: E402

DEV = "cuda:0"


def _group(world, cap):
    g = [_dist.PeerExchange(DEV, world, r, cap) for r in range(world)]
    _dist.PeerExchange.connect_local(g)
    return g


@pytest.mark.timeout(60, method="thread")
@pytest.mark.parametrize("world,count,wait", [(2, 1 << 20, "memop"), (2, 1027, "memop"), (3, 77777, "memop"), (4, 4, "memop"),
                                              (8, 3, "memop"), (8, 866_310, "memop"), (2, 1 << 20, "spin"), (4, 77777, "spin")])
def test_allreduce_matches_rank_ordered_sum(world, count, wait, monkeypatch):
    """(The emulated ranks' streams must not share a hardware queue -- a rank waiting for a peer queued behind it would
    never be released: conftest.py raises CUDA_DEVICE_MAX_CONNECTIONS before the context exists.)"""
    monkeypatch.setenv("A2CB_P2P_WAIT", wait)
    g = _group(world, count)
    streams = [torch.cuda.Stream(DEV) for _ in range(world)]
    gen = torch.Generator().manual_seed(world * 1000 + count % 997)
    for rep in range(3):                                      # consecutive exchanges reuse inbox / result / counters
        xs = [torch.randn(count + 4, generator=gen).to(DEV) for _ in range(world)]
        tails = [x[count:].clone() for x in xs]
        want = xs[0][:count].clone()
        for x in xs[1:]:
            want += x[:count]                                 # fp32, rank order: the kernel's order of additions
        torch.cuda.synchronize()
        for r in range(world):
            g[r].all_reduce_sum(xs[r][:count], stream=streams[r])
        torch.cuda.synchronize()
        for r in range(world):
            assert torch.equal(xs[r][:count], want), (rep, r)
            assert torch.equal(xs[r][count:], tails[r])
    for x in g:
        x.close()


@pytest.mark.timeout(60, method="thread")
def test_untouched_tail_and_argument_checks():
    world, count = 2, 1001
    g = _group(world, count)
    xs = [torch.full((count + 8,), float(r + 1), device=DEV) for r in range(world)]
    s = [torch.cuda.Stream(DEV) for _ in range(world)]
    for r in range(world):
        g[r].all_reduce_sum(xs[r][:count], stream=s[r])
    torch.cuda.synchronize()
    for r in range(world):
        assert bool((xs[r][:count] == 3.0).all()) and bool((xs[r][count:] == float(r + 1)).all())
    with pytest.raises(_native.NativeError):
        g[0].all_reduce_sum(torch.zeros(g[0].capacity + 4096, device=DEV))        # larger than the exchange buffer
    with pytest.raises(_native.NativeError):
        g[0].all_reduce_sum(torch.zeros(64, device=DEV)[1:33])                    # not 16-byte aligned
    lone = _dist.PeerExchange(DEV, 2, 0, 64)
    with pytest.raises(_native.NativeError):
        lone.all_reduce_sum(torch.zeros(16, device=DEV))                          # peers never mapped
    lone.close()
    for x in g:
        x.close()


@pytest.mark.timeout(60, method="thread")
def test_overlaps_with_a_kernel_that_fills_the_gpu():
    """The exchange must make progress next to work that keeps every SM busy on another stream (its kernels use no
    shared memory, its waits no SM): here a long matmul chain on the default stream."""
    world, count = 4, 1 << 18
    g = _group(world, count)
    a = torch.randn(4096, 4096, device=DEV)
    xs = [torch.randn(count, device=DEV) for _ in range(world)]
    want = xs[0].clone()
    for x in xs[1:]:
        want += x
    s = [torch.cuda.Stream(DEV) for _ in range(world)]
    torch.cuda.synchronize()
    b = a
    for _ in range(20):
        b = (b @ a) * 1e-2
    for r in range(world):
        g[r].all_reduce_sum(xs[r], stream=s[r])
    torch.cuda.synchronize()
    for r in range(world):
        assert torch.equal(xs[r], want)
    assert np.isfinite(float(b.abs().max()))
    for x in g:
        x.close()
