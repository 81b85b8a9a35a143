"""TMA-fed tensor-core engine (csrc/tcx.cu) against fp64 references: forward / data-gradient boxes (4-D, parity
view, out-of-bounds zero fill) and MN-major weight gradients.  Tolerance 2e-5 of the result's max magnitude
(3xTF32 products are ~1e-6; the remainder is the tensor core's truncating fp32 accumulation over a chain)."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from a2c_ppo_acktr import _native, _tcx as X  # noqa: E402

DEV = "cuda:0"
TOL = 2e-5


def lib():
    return _native.lib()


def st():
    return torch.cuda.current_stream().cuda_stream


def lo_plane(x):
    lo = torch.empty_like(x)
    _native.check(lib().a2cb_tcx_lo_plane(lo.data_ptr(), x.data_ptr(), x.numel(), st()), "lo")
    return lo


def tab(values):
    return torch.tensor(values, dtype=torch.int32, device=DEV)


def prep(src, N, K, ntab, ktab):
    img = torch.empty(N, K, device=DEV)
    img_lo = torch.empty(N, K, device=DEV)
    nt, kt = tab(ntab), tab(ktab)
    _native.check(lib().a2cb_tcx_prep_weights(img.data_ptr(), img_lo.data_ptr(), src.data_ptr(), N, K, nt.data_ptr(),
                                              kt.data_ptr(), st()), "prep")
    torch.cuda.synchronize()
    return img, img_lo


def rel(got, want):
    want = want.to(torch.float64)
    return float((got.double().cpu() - want.cpu()).abs().max() / want.abs().max())


def fwd_desc(A, A_lo, dims, strides, box, img, img_lo, kbs, tiles, out, out_lo, o_stride, o_ext, N, alpha=1.0,
             bias=None, act=0, mask=None, chain=0, o_base=0):
    d = X.FwdDesc()
    d.A = X.view(A.data_ptr(), A_lo.data_ptr() if A_lo is not None else None, dims, strides, box)
    d.B, d.B_lo = img.data_ptr(), img_lo.data_ptr()
    d.b_rows, d.b_cols = img.shape
    d.n_kb, d.chain = len(kbs), chain
    for i, (a, b) in enumerate(kbs):
        for j in range(4):
            d.kb_a[i][j] = a[j]
        d.kb_b[i] = b
    d.tiles = tiles
    d.out, d.out_lo = out.data_ptr(), (out_lo.data_ptr() if out_lo is not None else None)
    d.mask = mask.data_ptr() if mask is not None else None
    d.bias = bias.data_ptr() if bias is not None else None
    for i in range(5):
        d.o_stride[i] = o_stride[i]
        d.o_ext[i] = o_ext[i]
    d.o_base, d.N, d.alpha, d.act = o_base, N, alpha, act
    return d


def run_fwd(d):
    _native.check(lib().a2cb_tcx_fwd(ctypes.byref(d), st()), "tcx_fwd")
    torch.cuda.synchronize()


def case_linear(with_lo=True, chain=0):
    g = torch.Generator().manual_seed(1)
    M, K, N = 300, 96, 40
    A = torch.randn(M, K, generator=g).to(DEV)
    W = torch.randn(N, K, generator=g).to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    img, img_lo = prep(W, N, K, [n * K for n in range(N)], list(range(K)))
    out = torch.full((M, N), float("nan"), device=DEV)
    out_lo = torch.zeros(M, N, device=DEV)
    kbs = [((32 * k, 0, 0, 0), 32 * k) for k in range(K // 32)]
    A_lo = lo_plane(A) if with_lo else None
    d = fwd_desc(A, A_lo, (K, M), (1, K), (32, 128), img, img_lo, kbs,
                 X.tiling((M + 127) // 128, 1, 128), out, out_lo, (0, N, 0, 0, 0), (0, M, 1, 1, 1), N, bias=b, act=1,
                 chain=chain)
    run_fwd(d)
    want = torch.relu(A.double() @ W.double().t() + b.double())
    e = rel(out, want)
    e2 = rel(out + out_lo, want)
    return e, e2


def case_conv3():
    g = torch.Generator().manual_seed(2)
    B = 7
    x = torch.randn(B, 9, 9, 64, generator=g).to(DEV)
    W = torch.randn(32, 64, 3, 3, generator=g).to(DEV) * 0.1
    # image [out][(ky, kx, c)] = W[out, c, ky, kx]
    img, img_lo = prep(W, 32, 576, [n * 576 for n in range(32)], [c * 9 + t for t in range(9) for c in range(64)])
    out = torch.full((B, 7, 7, 32), float("nan"), device=DEV)
    kbs = [((c0, kx, ky, 0), (ky * 3 + kx) * 64 + c0) for ky in range(3) for kx in range(3) for c0 in (0, 32)]
    k = 5
    x_lo = lo_plane(x)
    d = fwd_desc(x, x_lo, (64, 9, 9, B), (1, 64, 576, 5184), (32, 7, 7, k), img, img_lo, kbs,
                 X.tiling((B + k - 1) // k, 3, k), out, None, (0, 32, 7 * 32, 49 * 32, 0), (0, 7, 7, B, 1), 32)
    run_fwd(d)
    want = F.conv2d(x.double().permute(0, 3, 1, 2).cpu(), W.double().cpu()).permute(0, 2, 3, 1)
    return rel(out, want)


def case_conv2_parity():
    g = torch.Generator().manual_seed(3)
    B = 5
    x = torch.randn(B, 20, 20, 32, generator=g).to(DEV)
    W = torch.randn(64, 32, 4, 4, generator=g).to(DEV) * 0.1
    img, img_lo = prep(W, 64, 512, [n * 512 for n in range(64)], [c * 16 + t for t in range(16) for c in range(32)])
    out = torch.full((B, 9, 9, 64), float("nan"), device=DEV)
    kbs = []
    for ky in range(4):
        for kx in range(4):
            kbs.append((((kx % 2) * 32, kx // 2, ky % 2, ky // 2), (ky * 4 + kx) * 32))
    k = 3
    x_lo = lo_plane(x)
    d = fwd_desc(x, x_lo, (64, 10, 2, 10, B), (1, 64, 640, 1280, 12800), (32, 9, 1, 9, k), img, img_lo, kbs,
                 X.tiling((B + k - 1) // k, 4, k), out, None, (0, 64, 0, 9 * 64, 81 * 64), (0, 9, 1, 9, B), 64)
    run_fwd(d)
    want = F.conv2d(x.double().permute(0, 3, 1, 2).cpu(), W.double().cpu(), stride=2).permute(0, 2, 3, 1)
    return rel(out, want)


def case_dgrad3():
    g = torch.Generator().manual_seed(4)
    B = 4
    dy = torch.randn(B, 7, 7, 32, generator=g).to(DEV)
    W = torch.randn(32, 64, 3, 3, generator=g).to(DEV) * 0.1
    mask = torch.randn(B, 9, 9, 64, generator=g).to(DEV)
    # image [ci][(ky, kx, co)] = W[co, ci, ky, kx]
    img, img_lo = prep(W, 64, 288, [ci * 9 for ci in range(64)], [co * 576 + t for t in range(9) for co in range(32)])
    out = torch.full((B, 9, 9, 64), float("nan"), device=DEV)
    kbs = [((0, -kx, -ky, 0), (ky * 3 + kx) * 32) for ky in range(3) for kx in range(3)]
    k = 3
    dy_lo = lo_plane(dy)
    d = fwd_desc(dy, dy_lo, (32, 7, 7, B), (1, 32, 224, 1568), (32, 9, 9, k), img, img_lo, kbs,
                 X.tiling((B + k - 1) // k, 3, k), out, None, (0, 64, 9 * 64, 81 * 64, 0), (0, 9, 9, B, 1), 64, mask=mask)
    run_fwd(d)
    want = F.conv_transpose2d(dy.double().permute(0, 3, 1, 2).cpu(), W.double().cpu()).permute(0, 2, 3, 1)
    want = want * (mask.cpu() > 0)
    return rel(out, want)


def trunc13(x):
    return (x.view(torch.int32) & -8192).view(torch.float32)


def rna13(x):
    return ((x.view(torch.int32) + 4096) & -8192).view(torch.float32)


def planes(x, variant):
    """(hi plane, lo plane) for the variants of the operand-split probe."""
    if variant == "raw+trunc_lo":
        return x, x - trunc13(x)
    if variant == "trunc+trunc_lo":
        return trunc13(x).contiguous(), x - trunc13(x)
    if variant == "raw+rna_lo":
        return x, x - rna13(x)
    if variant == "rna+rna_lo":
        return rna13(x).contiguous(), x - rna13(x)
    if variant == "rna+rna_lo_rounded":
        r = x - rna13(x)
        return rna13(x).contiguous(), rna13(r).contiguous()
    raise ValueError(variant)


def wgrad_linear(R=64, chain=2, variant=None, ints_a=False, ints_b=False):
    g = torch.Generator().manual_seed(5)
    M, Ka, N = 1000, 256, 64
    Xa = torch.randn(M, Ka, generator=g).to(DEV)
    dY = torch.randn(M, N, generator=g).to(DEV)
    if ints_a:
        Xa = torch.round(Xa * 4)
    if ints_b:
        dY = torch.round(dY * 4)
    d = X.WgradDesc()
    if variant is None:
        keep = (Xa, lo_plane(Xa), dY, lo_plane(dY))
    else:
        keep = planes(Xa, variant) + planes(dY, variant)
    keep = tuple(t.contiguous() for t in keep)
    d.A = X.view(keep[0].data_ptr(), keep[1].data_ptr(), (Ka, M), (1, Ka), (32, R))
    d.Bm = X.view(keep[2].data_ptr(), keep[3].data_ptr(), (N, M), (1, N), (32, R))
    d.G, d.nbb = 1, N // 32
    for j in range(4):
        d.ga[0][j][0] = 32 * j
    for j in range(N // 32):
        d.gb[j][0] = 32 * j
    n_rt = (M + R - 1) // R
    d.rows = X.tiling(n_rt, 1, R)
    per = 6
    d.splits, d.chain = (n_rt + per - 1) // per, chain
    d.ot_a, d.ot_b, d.ot_a_step, d.ot_b_step = Ka // 128, 1, 128, 0
    partial = torch.zeros(d.splits, N * Ka, device=DEV)
    row_off = torch.arange(Ka, dtype=torch.int32).to(DEV)
    d.partial, d.split_stride, d.row_off, d.col_stride, d.n_cols, d.alpha = partial.data_ptr(), N * Ka, row_off.data_ptr(), Ka, N, 1.0
    _native.check(lib().a2cb_tcx_wgrad(ctypes.byref(d), st()), "tcx_wgrad")
    torch.cuda.synchronize()
    got = partial.sum(0).view(N, Ka)
    want = dY.double().t() @ Xa.double()
    return rel(got, want)


def wgrad_conv3():
    g = torch.Generator().manual_seed(6)
    B = 6
    a2 = torch.randn(B, 9, 9, 64, generator=g).to(DEV)
    dy = torch.randn(B, 7, 7, 32, generator=g).to(DEV)
    d = X.WgradDesc()
    a2_lo, dy_lo = lo_plane(a2), lo_plane(dy)
    d.A = X.view(a2.data_ptr(), a2_lo.data_ptr(), (64, 9, 9, B), (1, 64, 576, 5184), (32, 7, 7, 1))
    d.Bm = X.view(dy.data_ptr(), dy_lo.data_ptr(), (32, 7, 7, B), (1, 32, 224, 1568), (32, 7, 7, 1))
    d.G, d.nbb = 5, 1
    offs = []
    for gi in range(5):
        for j in range(4):
            tap = 2 * gi + j // 2
            if tap < 9:
                d.ga[gi][j][0], d.ga[gi][j][1], d.ga[gi][j][2], d.ga[gi][j][3] = (j % 2) * 32, tap % 3, tap // 3, 0
            else:
                d.ga[gi][j][0], d.ga[gi][j][1], d.ga[gi][j][2], d.ga[gi][j][3] = 0, 0, 0, X.OOB
            for c in range(32):
                offs.append(((j % 2) * 32 + c) * 9 + tap if tap < 9 else -1)
    d.rows = X.tiling(B, 3, 1)
    d.splits, d.chain = 2, 2
    d.ot_a, d.ot_b, d.ot_a_step, d.ot_b_step = 1, 1, 0, 0
    partial = torch.zeros(d.splits, 32 * 576, device=DEV)
    row_off = torch.tensor(offs, dtype=torch.int32).to(DEV)
    d.partial, d.split_stride, d.row_off, d.col_stride, d.n_cols, d.alpha = partial.data_ptr(), 32 * 576, row_off.data_ptr(), 576, 32, 1.0
    _native.check(lib().a2cb_tcx_wgrad(ctypes.byref(d), st()), "tcx_wgrad")
    torch.cuda.synchronize()
    got = partial.sum(0).view(32, 64, 3, 3)
    xin = a2.double().permute(0, 3, 1, 2).cpu().requires_grad_(False)
    W = torch.zeros(32, 64, 3, 3, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(xin, W)
    (y * dy.double().permute(0, 3, 1, 2).cpu()).sum().backward()
    return rel(got, W.grad)




def test_fwd_linear_bias_relu_and_lo_plane():
    e, _ = case_linear()
    assert e <= TOL
    e, _ = case_linear(chain=1)
    assert e <= TOL


def test_truncation_semantics_lo_plane_matters():
    """kind::tf32 ignores the 13 low mantissa bits: without the lo plane the error is the TF32 rounding level."""
    e, _ = case_linear(with_lo=False)
    assert 5e-5 <= e <= 2e-3


def test_fwd_conv3_box():
    assert case_conv3() <= TOL


def test_fwd_conv2_parity_view():
    assert case_conv2_parity() <= TOL


def test_dgrad_conv3_zero_fill_and_mask():
    assert case_dgrad3() <= TOL


@pytest.mark.parametrize("R", [64, 44])
def test_wgrad_linear(R):
    assert wgrad_linear(R=R) <= TOL


def test_wgrad_exact_on_small_integers():
    assert wgrad_linear(ints_a=True, ints_b=True) == 0.0


def test_wgrad_conv3_groups_and_oob_boxes():
    assert wgrad_conv3() <= TOL


def test_s2d_and_colsum():
    g = torch.Generator().manual_seed(9)
    frames = torch.randint(0, 256, (6, 4, 84, 84), generator=g, dtype=torch.uint8).to(DEV)
    idx = torch.tensor([4, 0, 5], device=DEV)
    out = torch.empty(3, 21, 21, 64, device=DEV)
    _native.check(lib().a2cb_tcx_s2d(out.data_ptr(), None, frames.data_ptr(), 1, idx.data_ptr(), 3, 4, 84, 84, 4, st()), "s2d")
    want = frames[idx].float().view(3, 4, 21, 4, 21, 4).permute(0, 2, 4, 1, 3, 5).reshape(3, 21, 21, 64)
    assert torch.equal(out, want)
    x = torch.randn(1000, 96, generator=g).to(DEV)
    blocks = lib().a2cb_tcx_colsum_blocks(1000)
    part = torch.zeros(blocks, 96, device=DEV)
    lo = torch.zeros_like(x)
    _native.check(lib().a2cb_tcx_colsum(part.data_ptr(), x.data_ptr(), lo.data_ptr(), 1000, 96, 96, st()), "colsum")
    assert rel(part.sum(0), x.double().sum(0)) <= 1e-6
    assert torch.equal(lo, lo_plane(x))
