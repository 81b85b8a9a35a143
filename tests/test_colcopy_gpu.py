"""a2cb_copy_columns (csrc/colcopy.cu): column copies between rollout-shaped tensors against torch indexing."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from a2c_ppo_acktr import _native  # noqa: E402

DEV = "cuda:0"


def _fields(T, N, g):
    return [torch.randint(0, 256, (T + 4, N, 84, 84), generator=g, dtype=torch.uint8).to(DEV),       # frames: 7,056 B
            torch.randint(0, 5, (T + 1, N), generator=g, dtype=torch.uint8).to(DEV),                 # stack counters: 1 B
            torch.randn(T + 1, N, 1, generator=g).to(DEV),                                           # masks: 4 B
            torch.randint(0, 6, (T, N, 1), generator=g).to(DEV),                                     # actions: 8 B
            torch.randn(T + 1, N, 512, generator=g).to(DEV)[:1],                                     # initial state: 2 KiB, 1 row
            torch.randn(T, N, 3, generator=g).to(DEV)]                                               # 12 B (byte loop)


@pytest.mark.parametrize("T,N,n", [(5, 9, 4), (16, 32, 32), (3, 4, 1)])
def test_pack_unpack_and_column_copy(T, N, n):
    g = torch.Generator().manual_seed(T * 100 + N)
    src = _fields(T, N, g)
    cols = torch.randperm(N, generator=g)[:n].to(DEV)
    widths = [(f[:, 0].numel() * f.element_size() + 15) // 16 * 16 for f in src]
    cap = n + 3
    rec = torch.zeros(cap, sum(widths), dtype=torch.uint8, device=DEV)

    def views(buf, like):
        out, off = [], 0
        for f, w in zip(like, widths):
            wb = f[0, 0].numel() * f.element_size()
            out.append(buf[:, off:off + f.shape[0] * wb].view(cap, f.shape[0], wb).transpose(0, 1))
            off += w
        return out
    # pack: record j = columns cols[j] of every field, environment-major
    _native.copy_columns(list(zip(views(rec, src), src)), None, cols, n)
    off = 0
    for f, w in zip(src, widths):
        want = f.index_select(1, cols).transpose(0, 1).reshape(n, -1).contiguous().view(torch.uint8)
        assert torch.equal(rec[:n, off:off + want.shape[1]], want)
        off += w
    assert int(rec[n:].abs().sum()) == 0                                                  # records beyond n untouched
    # unpack into other columns of wider tensors
    M = N + 7
    dst = [torch.zeros((f.shape[0], M) + tuple(f.shape[2:]), dtype=f.dtype, device=DEV) for f in src]
    into = (torch.randperm(M, generator=g)[:n]).to(DEV)
    _native.copy_columns(list(zip(dst, views(rec, dst))), into, None, n)
    for d, f in zip(dst, src):
        assert torch.equal(d.index_select(1, into), f.index_select(1, cols))
        rest = torch.ones(M, dtype=torch.bool, device=DEV)
        rest[into] = False
        assert int(d[:, rest].abs().sum()) == 0
    # column -> column (the same list on both sides: uploaded frames into the shadow rollout)
    dst2 = [torch.zeros((f.shape[0], M) + tuple(f.shape[2:]), dtype=f.dtype, device=DEV) for f in src]
    _native.copy_columns(list(zip(dst2, src)), cols, cols, n)
    for d, f in zip(dst2, src):
        assert torch.equal(d.index_select(1, cols), f.index_select(1, cols))


def test_argument_checks():
    a = torch.zeros(4, 3, 8, device=DEV)
    with pytest.raises(_native.NativeError):
        _native.copy_columns([(a, torch.zeros(5, 3, 8, device=DEV))], None, None, 2)      # different rows
    with pytest.raises(_native.NativeError):
        _native.copy_columns([(a, torch.zeros(4, 3, 9, device=DEV))], None, None, 2)      # different element size
    with pytest.raises(_native.NativeError):
        _native.copy_columns([(a, a.clone())], torch.zeros(1, dtype=torch.int64, device=DEV), None, 2)   # short list
    _native.copy_columns([(a, a.clone())], None, None, 0)                                  # nothing to do
