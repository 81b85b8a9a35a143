"""Host-side planner for the affine-indexed GEMM engine (csrc/gemm.cu).

Every dense op of the policy networks -- convolution forward, its gather-form
transposed convolution (dgrad) over a zero-haloed gradient buffer, its weight
gradient, Linear forward / dgrad / wgrad -- is expressed as ONE descriptor of

    C[rowC(i) + colC(j)] = epi( alpha * sum_r A[rowA(i) + redA(r)] * B[redB(r) + colB(j)] )

with 3-digit affine index maps.  This module only does integer bookkeeping (no
tensor math): it is pure host logic, and `emulate()` evaluates a descriptor with
numpy so that each plan can be checked on the CPU against torch's conv2d / linear
and their autograd gradients (tests/test_plans_cpu.py).  The CUDA engine is then
validated once, generically, against `emulate()` on the GPU.

Layouts: activations are NHWC ([B, H, W, C] row-major) except the network input,
which stays in the rollout buffer's NCHW layout; weights and biases stay in the
reference's PyTorch layouts ([out, in, kh, kw], [out, in]) inside the flat
parameter buffer, so `state_dict()` is interchangeable with the reference.
"""
import ctypes

import numpy as np

c_i32, c_i64, c_f32, c_vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p


class Affine3(ctypes.Structure):
    _fields_ = [("d1", c_i64), ("d2", c_i64), ("s0", c_i64), ("s1", c_i64), ("s2", c_i64)]


class GemmDesc(ctypes.Structure):
    """Mirror of a2cb_gemm_t (include/a2cb.h)."""
    _fields_ = [
        ("A", c_vp), ("B", c_vp), ("C", c_vp), ("C_ones", c_vp), ("bias", c_vp), ("mask", c_vp),
        ("gather_row", c_vp), ("gather_red", c_vp),
        ("M", c_i64), ("N", c_i64), ("K", c_i64),
        ("rowA", Affine3), ("rowC", Affine3), ("rowM", Affine3),
        ("redA", Affine3), ("redB", Affine3),
        ("colB", Affine3), ("colC", Affine3), ("colM", Affine3),
        ("a_dtype", c_i32), ("ones_row", c_i32), ("ones_stride", c_i64),
        ("bias_mode", c_i32), ("act", c_i32), ("mask_mode", c_i32), ("accumulate", c_i32),
        ("alpha", c_f32), ("batch", c_i32),
        ("batch_offA", c_i64 * 4), ("batch_offB", c_i64 * 4), ("batch_offC", c_i64 * 4), ("batch_offM", c_i64 * 4),
        ("splits", c_i32), ("vecA", c_i32), ("split_stride", c_i64), ("vecB", c_i32), ("tile", c_i32),
        ("tc_chunk", c_i32), ("sym", c_i32), ("B_image", c_vp),
    ]


ACT_NONE, ACT_RELU, ACT_TANH = 0, 1, 2
# Weight gradients sum long, heavily cancelling series: keep their tensor-core accumulation chains short
# (2 k-blocks = 64 reduction indices per TMEM chain, then an fp32 round-to-nearest add in registers).
WGRAD_CHUNK = 2


def aff(d1, d2, s0, s1, s2):
    return (int(d1), int(d2), int(s0), int(s1), int(s2))


def plain(stride):
    """off(i) = i * stride."""
    return aff(1, 1, stride, 0, 0)


def aff_eval(a, idx):
    d1, d2, s0, s1, s2 = a
    idx = np.asarray(idx, dtype=np.int64)
    q0 = idx // d1
    rem = idx - q0 * d1
    q1 = rem // d2
    q2 = rem - q1 * d2
    return q0, q1 * s1 + q2 * s2, s0


class Plan(object):
    """A descriptor with SYMBOLIC operands: buffer names + element offsets, resolved to device
    pointers by the engine (or to numpy arrays by `emulate`)."""

    def __init__(self, name, A, B, C, M, N, K, rowA, redA, redB, colB, rowC, colC, **kw):
        self.name = name
        self.A, self.B, self.C = A, B, C            # (buffer name, element offset)
        self.M, self.N, self.K = int(M), int(N), int(K)
        self.rowA, self.redA, self.redB, self.colB, self.rowC, self.colC = rowA, redA, redB, colB, rowC, colC
        self.rowM = kw.pop("rowM", plain(0))
        self.colM = kw.pop("colM", plain(0))
        self.mask = kw.pop("mask", None)            # (buffer, offset)
        self.mask_mode = kw.pop("mask_mode", 0)
        self.bias = kw.pop("bias", None)
        self.bias_mode = kw.pop("bias_mode", 0)
        self.act = kw.pop("act", ACT_NONE)
        self.ones = kw.pop("ones", None)            # (buffer, offset) -> C_ones
        self.ones_stride = kw.pop("ones_stride", 1)
        self.a_u8 = kw.pop("a_u8", False)           # False/0 raw fp32, True/1 uint8/255, 2 fp32/255
        self.gather_row = kw.pop("gather_row", False)   # take q0 of rowA through the minibatch index list
        self.gather_red = kw.pop("gather_red", False)
        self.batch = kw.pop("batch", 1)
        self.batch_offA = kw.pop("batch_offA", [0] * 4)
        self.batch_offB = kw.pop("batch_offB", [0] * 4)
        self.batch_offC = kw.pop("batch_offC", [0] * 4)
        self.batch_offM = kw.pop("batch_offM", [0] * 4)
        self.splits = kw.pop("splits", 1)
        self.split_stride = kw.pop("split_stride", 0)
        self.split_dst = kw.pop("split_dst", None)  # where reduce_splits folds the partials: (buffer, offset, count)
        self.split_bias = kw.pop("split_bias", None)  # epilogue applied by the fold: bias[e % N] ...
        self.split_act = kw.pop("split_act", ACT_NONE)  # ... and activation
        self.split_beta = kw.pop("split_beta", 1.0)     # fold computes beta * dst + sum when accumulate
        self.vecA = kw.pop("vecA", 0)
        self.vecB = kw.pop("vecB", 0)
        self.alpha = kw.pop("alpha", 1.0)
        self.accumulate = kw.pop("accumulate", 0)
        self.tile = kw.pop("tile", 0)
        self.tc_chunk = kw.pop("tc_chunk", 0)
        self.sym = kw.pop("sym", 0)                 # symmetric product: the engine computes the upper tiles only
        self.b_image = kw.pop("b_image", None)      # (buffer, offset) of the pre-tiled hi/lo image of B
        self.flops = 2 * self.M * self.N * self.K * self.batch
        assert not kw, kw


def emulate(plan, bufs, gather=None):
    """Evaluates `plan` on numpy buffers (dict name -> 1-D array).  fp64 accumulation."""
    A = bufs[plan.A[0]]
    B = bufs[plan.B[0]]
    C = bufs[plan.C[0]]
    i = np.arange(plan.M)
    j = np.arange(plan.N)
    r = np.arange(plan.K)

    def offs(a, idx, g=None):
        q0, rest, s0 = aff_eval(a, idx)
        if g is not None:
            q0 = np.asarray(g, dtype=np.int64)[q0]
        return q0 * s0 + rest

    ra = offs(plan.rowA, i, gather if plan.gather_row else None)
    ka = offs(plan.redA, r, gather if plan.gather_red else None)
    kb = offs(plan.redB, r)
    cb = offs(plan.colB, j)
    rc = offs(plan.rowC, i)
    cc = offs(plan.colC, j)
    for z in range(plan.batch):
        Am = A[plan.A[1] + plan.batch_offA[z] + ra[:, None] + ka[None, :]].astype(np.float64)
        if plan.a_u8:                               # both frame encodings: value / 255 in fp32
            Am = (A[plan.A[1] + plan.batch_offA[z] + ra[:, None] + ka[None, :]].astype(np.float32)
                  / np.float32(255.0)).astype(np.float64)
        Bm = B[plan.B[1] + plan.batch_offB[z] + kb[:, None] + cb[None, :]].astype(np.float64)
        out = plan.alpha * (Am @ Bm)
        if plan.ones is not None and z == 0:
            ob = bufs[plan.ones[0]]
            ob[plan.ones[1] + j * plan.ones_stride] = (plan.alpha * Bm.sum(0)).astype(ob.dtype)
        if plan.bias_mode == 1:
            out = out + bufs[plan.bias[0]][plan.bias[1] + j][None, :]
        elif plan.bias_mode == 2:
            out = out + bufs[plan.bias[0]][plan.bias[1] + i][:, None]
        if plan.act == ACT_RELU:
            out = np.maximum(out, 0)
        elif plan.act == ACT_TANH:
            out = np.tanh(out)
        if plan.mask_mode:
            mk = bufs[plan.mask[0]]
            mv = mk[plan.mask[1] + plan.batch_offM[z] + offs(plan.rowM, i)[:, None] + offs(plan.colM, j)[None, :]]
            out = out * (mv > 0) if plan.mask_mode == 1 else out * (1.0 - mv.astype(np.float64) ** 2)
        dst = plan.C[1] + plan.batch_offC[z] + rc[:, None] + cc[None, :]
        if plan.accumulate:
            C[dst] = C[dst] + out.astype(C.dtype)
        else:
            C[dst] = out.astype(C.dtype)


# ---------------------------------------------------------------------------
# layer geometry
# ---------------------------------------------------------------------------
class Conv(object):
    """A square, unpadded, strided convolution (the reference uses no padding: model.py:177-179).

    in_layout: 'nchw' (network input, rollout-buffer layout) or 'nhwc'.
    The output activation is NHWC [B, OH, OW, Cout]; the gradient w.r.t. the output lives in a
    zero-haloed NHWC buffer [B, OH + 2P, OW + 2P, Cout], P = halo needed by THIS layer's dgrad.
    """

    def __init__(self, name, cin, h, w, k, s, cout, in_layout="nhwc", needs_dgrad=True):
        self.name, self.cin, self.h, self.w, self.k, self.s, self.cout = name, cin, h, w, k, s, cout
        self.in_layout = in_layout
        self.oh, self.ow = (h - k) // s + 1, (w - k) // s + 1
        self.needs_dgrad = needs_dgrad
        if needs_dgrad:
            assert k % s == 0 and h % s == 0 and w % s == 0, "gather-form dgrad needs k, H, W divisible by stride"
        self.halo = (k // s - 1) if needs_dgrad else 0
        self.K = cin * k * k
        self.in_row = cin * h * w                                  # elements per image
        self.out_row = self.oh * self.ow * cout
        self.ph, self.pw = self.oh + 2 * self.halo, self.ow + 2 * self.halo
        self.gout_row = self.ph * self.pw * cout                   # haloed gradient image
        self.gout_base = (self.halo * self.pw + self.halo) * cout  # offset of (oy, ox) = (0, 0)

    def vec_ok(self):
        """True when 4 consecutive K indices are contiguous in the input and every patch origin
        is 4-element aligned (16 B for fp32, 4 B for uint8)."""
        if self.in_layout == "nchw":
            return self.k % 4 == 0 and self.s % 4 == 0 and self.w % 4 == 0 and (self.h * self.w) % 4 == 0
        return self.cin % 4 == 0

    # K index orders: nchw -> (c, ky, kx) [== PyTorch weight order]; nhwc -> (ky, kx, c)
    def red_in(self):
        k, c, w = self.k, self.cin, self.w
        if self.in_layout == "nchw":
            return aff(k * k, k, self.h * w, w, 1)
        return aff(k * c, c, w * c, c, 1)

    def red_weight(self):
        """offset inside one output channel's [cin, k, k] weight block, as a function of the K index."""
        k, c = self.k, self.cin
        if self.in_layout == "nchw":
            return plain(1)
        return aff(k * c, c, k, 1, k * k)

    def row_in(self):
        """patch origin of output pixel m = (b, oy, ox) in the input."""
        s = self.s
        if self.in_layout == "nchw":
            return aff(self.oh * self.ow, self.ow, self.in_row, s * self.w, s)
        return aff(self.oh * self.ow, self.ow, self.in_row, s * self.w * self.cin, s * self.cin)

    def row_gout(self):
        """output pixel m = (b, oy, ox) in the haloed gradient buffer (add gout_base)."""
        return aff(self.oh * self.ow, self.ow, self.gout_row, self.pw * self.cout, self.cout)


def conv_forward(cv, B, src, w, b, dst, act=ACT_RELU, a_u8=False, gather=False, out_layout="nhwc"):
    """out_layout 'nchw' writes [B, Cout, OH, OW] (the order the reference's Flatten() feeds main.7)."""
    p = cv.oh * cv.ow
    if out_layout == "nhwc":
        rowC, colC = plain(cv.cout), plain(1)
    else:
        rowC, colC = aff(p, 1, cv.cout * p, 1, 0), plain(p)
    wk = cv.red_weight()
    return Plan(cv.name + ".fwd", src, w, dst, B * p, cv.cout, cv.K,
                rowA=cv.row_in(), redA=cv.red_in(), redB=wk, colB=plain(cv.K),
                rowC=rowC, colC=colC, bias=b, bias_mode=1, act=act, a_u8=a_u8,
                gather_row=gather, vecA=1 if cv.vec_ok() else 0,
                vecB=2 if (wk == plain(1) and cv.K % 4 == 0) else 0)


def conv_wgrad(cv, B, src, gout, dw, db, a_u8=False, gather=False, splits=1, partial=None):
    """dW[n, (c,ky,kx)] = sum_m gout(m, n) * patch(m, k);  db[n] = sum_m gout(m, n)  (ones row).
    rows i = K index, cols j = n, reduction over m = (b, oy, ox)."""
    k, c = cv.k, cv.cin
    row_w = plain(1) if cv.in_layout == "nchw" else aff(k * c, c, k, 1, k * k)
    vecA = 2 if cv.vec_ok() else 0
    kw = {}
    C, ones = dw, db
    if splits > 1:
        span = cv.cout * cv.K + (cv.cout if db is not None else 0)
        C = (partial[0], partial[1])
        ones = (partial[0], partial[1] + cv.cout * cv.K) if db is not None else None
        kw = dict(splits=splits, split_stride=span, split_dst=(dw[0], dw[1], span))
        assert db is None or (db[1] == dw[1] + cv.cout * cv.K and db[0] == dw[0]), \
            "bias must follow the weight in the flat buffer"
    return Plan(cv.name + ".wgrad", src, (gout[0], gout[1] + cv.gout_base), C, cv.K, cv.cout, B * cv.oh * cv.ow,
                rowA=cv.red_in(), redA=cv.row_in(), redB=cv.row_gout(), colB=plain(1),
                rowC=row_w, colC=plain(cv.K), ones=ones, ones_stride=1, a_u8=a_u8, gather_red=gather,
                vecA=vecA, vecB=1 if cv.cout % 4 == 0 else 0, tc_chunk=WGRAD_CHUNK, **kw)


def conv_dgrad(cv, B, gout, w, gin, gin_geom, mask):
    """Input gradient of `cv` in gather form, one GEMM per stride-parity class (batch = s*s):

        gin[b, s*jy+py, s*jx+px, c] = relu'(.) * sum_{a,bb,n} gout[b, jy-a, jx-bb, n] * W[n, c, py+s*a, px+s*bb]

    gout is haloed by cv.halo so no bounds checks are needed.  `gin_geom` = (row_stride, pitch_w,
    base) of the destination buffer (itself the haloed gradient buffer of the previous conv, or a
    plain NHWC buffer); `mask` is the previous layer's post-ReLU activation (plain NHWC)."""
    assert cv.in_layout == "nhwc" and cv.needs_dgrad
    s, k, c, n = cv.s, cv.k, cv.cin, cv.cout
    t = k // s
    jh, jw = cv.h // s, cv.w // s
    g_row, g_pw, g_base = gin_geom
    offB, offC, offM = [0] * 4, [0] * 4, [0] * 4
    assert s * s <= 4
    for py in range(s):
        for px in range(s):
            z = py * s + px
            offB[z] = py * k + px
            offC[z] = (py * g_pw + px) * c
            offM[z] = (py * cv.w + px) * c
    return Plan(cv.name + ".dgrad", (gout[0], gout[1] + cv.gout_base), w, (gin[0], gin[1] + g_base),
                B * jh * jw, c, t * t * n,
                rowA=aff(jh * jw, jw, cv.gout_row, cv.pw * n, n),
                redA=aff(t * n, n, -cv.pw * n, -n, 1),
                redB=aff(t * n, n, s * k, s, c * k * k), colB=plain(k * k),
                rowC=aff(jh * jw, jw, g_row, s * g_pw * c, s * c), colC=plain(1),
                rowM=aff(jh * jw, jw, cv.in_row, s * cv.w * c, s * c), colM=plain(1),
                mask=mask, mask_mode=1, batch=s * s, batch_offB=offB, batch_offC=offC, batch_offM=offM,
                vecA=1 if n % 4 == 0 else 0)


class Linear(object):
    """y = x W^T + b with W in the reference's [out, in] layout.  `in_perm` optionally maps my
    K order to the weight's column order (NHWC-flattened conv features vs the reference's
    Flatten() of NCHW, model.py:179-180)."""

    def __init__(self, name, d_in, d_out, in_perm=None):
        self.name, self.d_in, self.d_out = name, d_in, d_out
        self.in_perm = in_perm or plain(1)


def linear_forward(ln, B, x, w, b, y, act=ACT_NONE, x_stride=None, y_stride=None, splits=1, partial=None):
    kw = {}
    C, bias, bias_mode, gact = y, b, (1 if b is not None else 0), act
    if splits > 1:
        assert (y_stride or ln.d_out) == ln.d_out, "split-K forward needs a dense [B, d_out] output"
        C, bias, bias_mode, gact = partial, None, 0, ACT_NONE
        kw = dict(splits=splits, split_stride=B * ln.d_out, split_dst=(y[0], y[1], B * ln.d_out),
                  split_bias=b, split_act=act)
    return Plan(ln.name + ".fwd", x, w, C, B, ln.d_out, ln.d_in,
                rowA=plain(x_stride or ln.d_in), redA=plain(1), redB=ln.in_perm, colB=plain(ln.d_in),
                rowC=plain(y_stride or ln.d_out), colC=plain(1), bias=bias, bias_mode=bias_mode,
                act=gact, vecA=1 if (ln.d_in % 4 == 0 and (x_stride or ln.d_in) % 4 == 0) else 0,
                vecB=2 if (ln.in_perm == plain(1) and ln.d_in % 4 == 0 and w[1] % 4 == 0) else 0, **kw)


def linear_dgrad(ln, B, gy, w, gx, mask=None, mask_mode=0, gy_stride=None, col_out=None, row_out=None,
                 gx_base=0, accumulate=0, mask_row=None, mask_col=None, splits=1, partial=None):
    """gx[b, k] = mask * sum_n gy[b, n] W[n, perm(k)].  col_out/row_out let the result land in a
    haloed NHWC gradient buffer.  splits > 1 (dense [B, d_in] output, no mask): split-K partials folded
    (and, with accumulate, added onto gx) by reduce_splits -- for the tiny per-time-step GRU GEMMs."""
    gs = gy_stride or ln.d_out
    kw = {}
    C = (gx[0], gx[1] + gx_base)
    if splits > 1:
        assert mask is None and row_out is None and col_out is None and gx_base == 0
        kw = dict(splits=splits, split_stride=B * ln.d_in, split_dst=(gx[0], gx[1], B * ln.d_in))
        C = partial
    return Plan(ln.name + ".dgrad", gy, w, C, B, ln.d_in, ln.d_out,
                rowA=plain(gs), redA=plain(1), redB=plain(ln.d_in), colB=ln.in_perm,
                rowC=row_out or plain(ln.d_in), colC=col_out or plain(1),
                rowM=mask_row or plain(ln.d_in), colM=mask_col or plain(1), mask=mask, mask_mode=mask_mode,
                accumulate=accumulate, vecA=1 if (ln.d_out % 4 == 0 and gs % 4 == 0) else 0,
                vecB=1 if (ln.in_perm == plain(1) and ln.d_in % 4 == 0) else 0, **kw)


def linear_wgrad(ln, B, x, gy, dw, db, x_stride=None, gy_stride=None, splits=1, partial=None, accumulate=0):
    """dW[n, perm(k)] = sum_b gy[b, n] x[b, k];  db[n] = sum_b gy[b, n].  rows i = k, cols j = n."""
    xs, gs = x_stride or ln.d_in, gy_stride or ln.d_out
    kw = {}
    C, ones = dw, db
    if splits > 1:
        span = ln.d_out * ln.d_in + ln.d_out
        C, ones = (partial[0], partial[1]), (partial[0], partial[1] + ln.d_out * ln.d_in)
        kw = dict(splits=splits, split_stride=span, split_dst=(dw[0], dw[1], span))
        assert db is not None and db[1] == dw[1] + ln.d_out * ln.d_in
    return Plan(ln.name + ".wgrad", x, gy, C, ln.d_in, ln.d_out, B,
                rowA=plain(1), redA=plain(xs), redB=plain(gs), colB=plain(1),
                rowC=ln.in_perm, colC=plain(ln.d_in), ones=ones if db is not None else None, ones_stride=1,
                vecA=2 if (ln.d_in % 4 == 0 and xs % 4 == 0) else 0,
                vecB=1 if (ln.d_out % 4 == 0 and gs % 4 == 0) else 0, accumulate=accumulate,
                tc_chunk=WGRAD_CHUNK, **kw)
