"""Host-side descriptors of the TMA-fed tensor-core engine (csrc/tcx.cu): ctypes mirrors of
`a2cb_tcx_*_t` (include/a2cb.h) and small builders for the views / tilings the CNN layers use.
Pure integer bookkeeping -- no tensor math happens here."""
import ctypes
import os

from . import _native
from . import _plans as P

c_i16, c_i32, c_i64, c_f32, c_vp = ctypes.c_int16, ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p

MAX_KB, MAX_G = 64, 8
OOB = 30000          # coordinate offset that pushes a box completely out of bounds (TMA zero-fills it)


class View(ctypes.Structure):
    _fields_ = [("x", c_vp), ("lo", c_vp), ("dim", c_i64 * 5), ("stride", c_i64 * 5), ("box", c_i32 * 5), ("pad_", c_i32)]


class Tiling(ctypes.Structure):
    _fields_ = [("n_tiles", c_i32), ("t_div", c_i32), ("tr_dim", c_i32), ("tr_step", c_i32), ("tq_dim", c_i32),
                ("tq_step", c_i32)]


class FwdDesc(ctypes.Structure):
    _fields_ = [("A", View), ("B", c_vp), ("B_lo", c_vp), ("b_rows", c_i64), ("b_cols", c_i64), ("n_kb", c_i32),
                ("chain", c_i32), ("kb_a", (c_i16 * 4) * MAX_KB), ("kb_b", c_i32 * MAX_KB), ("tiles", Tiling),
                ("out", c_vp), ("out_lo", c_vp), ("mask", c_vp), ("bias", c_vp), ("o_stride", c_i64 * 5),
                ("o_base", c_i64), ("o_ext", c_i32 * 5), ("N", c_i32), ("alpha", c_f32), ("act", c_i32),
                ("o_cb", c_i32), ("gather_dim", c_i32), ("o_cb_off", c_i64 * 4), ("gather_idx", c_vp), ("colsum_part", c_vp),
                ("B3", c_vp), ("bf16", c_i32), ("patch", c_i32)]


class WgradDesc(ctypes.Structure):
    _fields_ = [("A", View), ("Bm", View), ("G", c_i32), ("nbb", c_i32), ("ga", ((c_i16 * 4) * 4) * MAX_G),
                ("gb", (c_i16 * 4) * 8), ("rows", Tiling), ("splits", c_i32), ("chain", c_i32), ("ot_a", c_i32),
                ("ot_b", c_i32), ("ot_a_step", c_i32), ("ot_b_step", c_i32), ("partial", c_vp), ("split_stride", c_i64),
                ("row_off", c_vp), ("col_stride", c_i64), ("n_cols", c_i32), ("alpha", c_f32), ("gather_idx", c_vp),
                ("gather_dim", c_i32), ("pad_", c_i32)]


class LoDesc(ctypes.Structure):
    _fields_ = [("lo", c_vp), ("x", c_vp), ("n", c_i64)]


class PrepDesc(ctypes.Structure):
    _fields_ = [("img", c_vp), ("img_lo", c_vp), ("src", c_vp), ("N", c_i64), ("K", c_i64), ("ntab", c_vp), ("ktab", c_vp)]


class S2dDesc(ctypes.Structure):
    _fields_ = [("out", c_vp), ("out_lo", c_vp), ("frames", c_vp), ("idx", c_vp), ("src_is_u8", c_i32), ("B", c_i32),
                ("C", c_i32), ("H", c_i32), ("W", c_i32), ("S", c_i32), ("out_b16", c_vp), ("scatter", c_i32), ("ring_n", c_i32),
                ("valid", c_vp)]


class ColsumDesc(ctypes.Structure):
    _fields_ = [("part", c_vp), ("x", c_vp), ("rows", c_i64), ("pitch", c_i64), ("C", c_i32), ("pad_", c_i32), ("lo", c_vp)]


class NarrowFwdDesc(ctypes.Structure):
    _fields_ = [("z", c_vp), ("x", c_vp), ("W", c_vp), ("bias", c_vp), ("rows", c_i64), ("K", c_i32), ("N", c_i32),
                ("z_stride", c_i32), ("pad_", c_i32)]


class NarrowDgradDesc(ctypes.Structure):
    _fields_ = [("gx", c_vp), ("gx_lo", c_vp), ("dz", c_vp), ("W", c_vp), ("mask", c_vp), ("rows", c_i64), ("K", c_i32),
                ("N", c_i32), ("dz_stride", c_i32), ("pad_", c_i32)]


class PrepJob(ctypes.Structure):
    _fields_ = [("img", c_vp), ("img_lo", c_vp), ("src", c_vp), ("ntab", c_vp), ("ktab", c_vp), ("N", c_i64), ("K", c_i64),
                ("first", c_i64), ("img3", c_vp), ("bf16", c_i64)]


class PrepMultiDesc(ctypes.Structure):
    _fields_ = [("jobs", c_vp), ("total", c_i64), ("n_jobs", c_i32), ("pad_", c_i32)]


class NarrowWgradDesc(ctypes.Structure):
    _fields_ = [("part", c_vp), ("x", c_vp), ("gy", c_vp), ("rows", c_i64), ("K", c_i32), ("N", c_i32),
                ("gy_stride", c_i32), ("pad_", c_i32)]


_native.register({
    "a2cb_tcx_narrow_fwd": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_vp]),
    "a2cb_tcx_narrow_dgrad": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_vp]),
    "a2cb_tcx_prep_multi": (c_i32, [c_vp, c_i32, c_i64, c_vp]),
    "a2cb_tcx_narrow_wgrad": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_vp]),
    "a2cb_tcx_narrow_wgrad_blocks": (c_i32, [c_i64]),
    "a2cb_tcx_fwd": (c_i32, [ctypes.POINTER(FwdDesc), c_vp]),
    "a2cb_tcx_wgrad": (c_i32, [ctypes.POINTER(WgradDesc), c_vp]),
    "a2cb_tcx_lo_plane": (c_i32, [c_vp, c_vp, c_i64, c_vp]),
    "a2cb_tcx_prep_weights": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp]),
    "a2cb_tcx_s2d": (c_i32, [c_vp, c_vp, c_vp, c_i32, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp]),
    "a2cb_tcx_s2d_b16": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i32, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp]),
    "a2cb_tcx_colsum": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_i32, c_i64, c_vp]),
    "a2cb_tcx_colsum_blocks": (c_i32, [c_i64]),
    "a2cb_tcx_sizeof": (c_i32, [c_i32]),
    "a2cb_tcx_set_knobs": (c_i32, [c_i32, c_i32, c_i32, c_i32]),
})


def view(x_ptr, lo_ptr, dims, strides, box, bf16=False):
    """5-D view, innermost first; shorter tuples are padded with unit dimensions."""
    v = View()
    v.x, v.lo = x_ptr, lo_ptr
    dims, strides, box = list(dims), list(strides), list(box)
    while len(dims) < 5:
        strides.append(strides[-1] * dims[-1])
        dims.append(1)
        box.append(1)
    assert strides[0] == 1 and box[0] == (64 if bf16 else 32)
    for d in range(5):
        v.dim[d], v.stride[d], v.box[d] = dims[d], strides[d], box[d]
    return v


def tiling(n_tiles, tr_dim, tr_step, t_div=None, tq_dim=4, tq_step=0):
    t = Tiling()
    t.n_tiles, t.t_div = n_tiles, (t_div if t_div else n_tiles)
    t.tr_dim, t.tr_step, t.tq_dim, t.tq_step = tr_dim, tr_step, tq_dim, tq_step
    return t


# ---------------------------------------------------------------------------------------------------
# op builders (symbolic: buffer references are (arena name, element offset) pairs resolved at program
# build time, exactly like the GEMM plans of _plans.py)
# ---------------------------------------------------------------------------------------------------
OP_TCX_FWD, OP_TCX_WGRAD, OP_TCX_LO, OP_TCX_PREP, OP_TCX_S2D, OP_TCX_COLSUM = 14, 15, 16, 17, 18, 19
OP_TCX_NARROW_WGRAD = 20
OP_TCX_PREP_MULTI = 21
OP_TCX_NARROW_FWD, OP_TCX_NARROW_DGRAD = 24, 25


class TcxOp(object):
    """One launch of the TMA-fed engine in a plan list."""

    def __init__(self, kind, make, name, runtime_idx=False, flops=0.0):
        self.kind, self.make, self.name, self.runtime_idx, self.flops = kind, make, name, runtime_idx, flops


def _itab(arena, name, values):
    import numpy as np
    return arena.const(np.ascontiguousarray(np.asarray(values, dtype=np.int32))).data_ptr()


def fwd_op(name, A, A_lo, dims, strides, box, img, img_rows, img_cols, kbs, tiles, out, out_lo, o_stride, o_ext, N,
           alpha=1.0, bias=None, act=0, mask=None, chain=4, o_base=0, flops=0.0, col_blocks=None, gather_dim=0,
           colsum_part=None, bf16=False, emu_A=None, patch=None):
    """bf16: A and the three-plane weight image (img, img_lo, img_3) are bf16, k-blocks are 64 channels wide
    (emu_A: fp32 tensor with the same values, read by the numpy emulator instead).
    patch = (rows, tile box): `box` is the input patch of a tile (tile pixels + tap halo) that is loaded once; the
    k-block offsets are tap shifts inside it; `rows` accumulator rows (patch pixels) are stored, `tile box` is the
    tile's own pixel box over the same row pitch (what the emulator enumerates)."""
    assert len(kbs) <= MAX_KB, (name, len(kbs))      # col_blocks: (block width, [offset per block]) -- a2cb_tcx_fwd_t.o_cb

    def make(arena):
        d = FwdDesc()
        d.A = view(arena.ptr(A), arena.ptr(A_lo), dims, strides, box, bf16)
        d.B, d.B_lo = arena.ptr((img, 0)), arena.ptr((img + "_lo", 0))
        if bf16:
            d.B3, d.bf16 = arena.ptr((img + "_3", 0)), 1
        d.b_rows, d.b_cols = img_rows, img_cols
        d.n_kb, d.chain = len(kbs), chain
        for i, (a, b) in enumerate(kbs):
            for j in range(4):
                d.kb_a[i][j] = a[j]
            d.kb_b[i] = b
        d.tiles = tiles
        d.out, d.out_lo = arena.ptr(out), arena.ptr(out_lo)
        d.mask, d.bias = arena.ptr(mask), arena.ptr(bias)
        os_, oe = list(o_stride) + [0] * (5 - len(o_stride)), list(o_ext) + [1] * (5 - len(o_ext))
        for i in range(5):
            d.o_stride[i], d.o_ext[i] = os_[i], oe[i]
        d.o_base, d.N, d.alpha, d.act = o_base, N, alpha, act
        if col_blocks is not None:
            d.o_cb = col_blocks[0]
            for i, o in enumerate(col_blocks[1]):
                d.o_cb_off[i] = o
        d.gather_dim = gather_dim
        d.gather_idx = 1 if gather_dim else None          # placeholder: a2cb_run_ops substitutes the runtime row list
        d.colsum_part = arena.ptr(colsum_part)
        if patch is not None:
            assert patch[1][1] == box[1] and patch[0] == patch[1][1] * patch[1][2] * patch[1][3]
            d.patch = patch[0]
        return d
    op = TcxOp(OP_TCX_FWD, make, name, flops=flops, runtime_idx=bool(gather_dim))
    op.spec = dict(A=emu_A if emu_A is not None else A, cw=64 if bf16 else 32,
                   dims=dims, strides=strides, box=patch[1] if patch is not None else box, img=img, img_rows=img_rows, img_cols=img_cols, kbs=kbs,
                   tiles=tiles, out=out, o_stride=o_stride, o_ext=o_ext, N=N, alpha=alpha, bias=bias, act=act, mask=mask,
                   o_base=o_base, col_blocks=col_blocks, gather_dim=gather_dim, colsum_part=colsum_part)
    return op


def wgrad_op(name, A, A_lo, a_dims, a_strides, a_box, Bm, Bm_lo, b_dims, b_strides, b_box, groups, b_boxes, rows, splits,
             chain, ot_a, ot_b, ot_a_step, ot_b_step, partial, split_stride, row_off, col_stride, n_cols, alpha=1.0,
             flops=0.0, gather_dim=0):
    """groups: per accumulator group, 4 coordinate-offset tuples (dims 0..3); b_boxes: per 32-channel box of Bm one
    tuple; row_off: list of ot_a * G * 128 element offsets (-1 = unused accumulator row)."""
    G = len(groups)
    assert 1 <= G <= MAX_G and all(len(g) == 4 for g in groups) and len(row_off) == ot_a * G * 128

    def make(arena):
        d = WgradDesc()
        d.A = view(arena.ptr(A), arena.ptr(A_lo), a_dims, a_strides, a_box)
        d.Bm = view(arena.ptr(Bm), arena.ptr(Bm_lo), b_dims, b_strides, b_box)
        d.G, d.nbb = G, len(b_boxes)
        for g, boxes in enumerate(groups):
            for j, c in enumerate(boxes):
                for k in range(4):
                    d.ga[g][j][k] = c[k]
        for j, c in enumerate(b_boxes):
            for k in range(4):
                d.gb[j][k] = c[k]
        d.rows, d.splits, d.chain = rows, splits, chain
        d.ot_a, d.ot_b, d.ot_a_step, d.ot_b_step = ot_a, ot_b, ot_a_step, ot_b_step
        d.partial, d.split_stride = arena.ptr(partial), split_stride
        d.row_off = _itab(arena, partial[0] + "_rowoff", row_off)
        d.col_stride, d.n_cols, d.alpha = col_stride, n_cols, alpha
        d.gather_dim = gather_dim
        d.gather_idx = 1 if gather_dim else None
        return d
    op = TcxOp(OP_TCX_WGRAD, make, name, flops=flops, runtime_idx=bool(gather_dim))
    op.spec = dict(A=A, a_dims=a_dims, a_strides=a_strides, a_box=a_box, Bm=Bm, b_dims=b_dims, b_strides=b_strides,
                   b_box=b_box, groups=groups, b_boxes=b_boxes, rows=rows, splits=splits, ot_a=ot_a, ot_b=ot_b,
                   ot_a_step=ot_a_step, ot_b_step=ot_b_step, partial=partial, split_stride=split_stride, row_off=row_off,
                   col_stride=col_stride, n_cols=n_cols, alpha=alpha, gather_dim=gather_dim)
    return op


def lo_op(name, x, lo, n):
    def make(arena):
        return LoDesc(arena.ptr(lo), arena.ptr(x), n)
    op = TcxOp(OP_TCX_LO, make, name)
    op.spec = dict(x=x, lo=lo, n=n)
    return op


def prep_op(name, img, src, N, K, ntab, ktab, bf16=False):
    """bf16: three bf16 planes (img, img_lo, img_3) with w = b1 + b2 + b3; only through the merged launch."""
    def make(arena):
        assert not bf16
        return PrepDesc(arena.ptr((img, 0)), arena.ptr((img + "_lo", 0)), arena.ptr(src), N, K,
                        _itab(arena, img + "_ntab", ntab), _itab(arena, img + "_ktab", ktab))
    op = TcxOp(OP_TCX_PREP, make, name)
    op.prep = (img, src, N, K, ntab, ktab)
    op.bf16 = bf16
    return op


def hoist_preps(lists, tag):
    """Weight images depend on the weights only: moves every weight-image op of the given plan lists to the front of
    the first one and merges them into one launch.  Returns the new lists."""
    is_prep = lambda o: isinstance(o, TcxOp) and o.kind == OP_TCX_PREP
    preps = [o for l in lists for o in l if is_prep(o)]
    rest = [[o for o in l if not is_prep(o)] for l in lists]
    rest[0] = merge_preps(preps, tag) + rest[0]
    return rest


def merge_preps(ops, tag):
    """Replaces runs of consecutive weight-image ops of a plan list by ONE launch each (a2cb_tcx_prep_multi)."""
    import numpy as np
    out, run = [], []

    def flush():
        if not run:
            return
        if len(run) == 1 and not run[0].bf16:
            out.extend(run)
        else:
            jobs_spec = [o.prep + (o.bf16,) for o in run]
            name = "%sprepjobs_%d" % (tag, len(out))

            def make(arena, jobs_spec=jobs_spec, name=name):
                arr = (PrepJob * len(jobs_spec))()
                first = 0
                for i, (img, src, N, K, ntab, ktab, b16) in enumerate(jobs_spec):
                    arr[i].img, arr[i].img_lo, arr[i].src = arena.ptr((img, 0)), arena.ptr((img + "_lo", 0)), arena.ptr(src)
                    if b16:
                        arr[i].img3, arr[i].bf16 = arena.ptr((img + "_3", 0)), 1
                    arr[i].ntab, arr[i].ktab = _itab(arena, img + "_ntab", ntab), _itab(arena, img + "_ktab", ktab)
                    arr[i].N, arr[i].K, arr[i].first = N, K, first
                    first += N * K
                raw = arena.const(np.frombuffer(bytes(arr), dtype=np.uint8))
                return PrepMultiDesc(raw.data_ptr(), first, len(jobs_spec), 0)
            out.append(TcxOp(OP_TCX_PREP_MULTI, make, "weight_image"))
        del run[:]
    for o in ops:
        if isinstance(o, TcxOp) and o.kind == OP_TCX_PREP:
            run.append(o)
        else:
            flush()
            out.append(o)
    flush()
    return out


def colsum_op(name, part, x, rows, C, pitch, lo=None):
    def make(arena):
        return ColsumDesc(arena.ptr(part), arena.ptr(x), rows, pitch, C, 0, arena.ptr(lo))
    op = TcxOp(OP_TCX_COLSUM, make, name)
    op.spec = dict(part=part, x=x, rows=rows, C=C, pitch=pitch, lo=lo, blocks=colsum_blocks(rows))
    return op


def fwd_grid(n_tiles, N):
    """CTAs of a forward-like launch (persistent kernel: min(work items, SMs)); mirrors launch_fwd in csrc/tcx.cu."""
    bn = 32 if N <= 32 else (64 if N <= 64 else 128)
    return int(min(n_tiles * (-(-N // bn)), 148))


def colsum_blocks(rows):
    # mirrors tcx_colsum_blocks (csrc/tcx_aux.cu)
    return int(min((rows + 15) // 16, 148 * 8))


def narrow_wgrad_blocks(rows):
    # mirrors tcx_narrow_wgrad_blocks (csrc/tcx_aux.cu)
    return int(min((rows + 63) // 64, 148 * 4))


def narrow_fwd_op(name, z, x, W, bias, rows, K, N, z_stride):
    def make(arena):
        return NarrowFwdDesc(arena.ptr(z), arena.ptr(x), arena.ptr(W), arena.ptr(bias), rows, K, N, z_stride, 0)
    op = TcxOp(OP_TCX_NARROW_FWD, make, name)
    op.spec = dict(z=z, x=x, W=W, bias=bias, rows=rows, K=K, N=N, z_stride=z_stride)
    return op


def narrow_dgrad_op(name, gx, gx_lo, dz, W, mask, rows, K, N, dz_stride):
    def make(arena):
        return NarrowDgradDesc(arena.ptr(gx), arena.ptr(gx_lo), arena.ptr(dz), arena.ptr(W), arena.ptr(mask), rows, K, N,
                               dz_stride, 0)
    op = TcxOp(OP_TCX_NARROW_DGRAD, make, name)
    op.spec = dict(gx=gx, gx_lo=gx_lo, dz=dz, W=W, mask=mask, rows=rows, K=K, N=N, dz_stride=dz_stride)
    return op


def narrow_wgrad_ops(nb, name, x, gy, rows, K, N, gy_stride, dst):
    """dW [N, K] and db [N] (contiguous at dst) of a narrow Linear layer: one streaming pass + one fold."""
    from ._engine import RawOp, ReduceDesc, OP_REDUCE
    blocks = narrow_wgrad_blocks(rows)
    span = N * K + N
    part = nb.buf("npart_" + name, blocks * span)

    def make(arena):
        return NarrowWgradDesc(arena.ptr(part), arena.ptr(x), arena.ptr(gy), rows, K, N, gy_stride, 0)
    nw = TcxOp(OP_TCX_NARROW_WGRAD, make, name + ".wgrad")
    nw.spec = dict(part=part, x=x, gy=gy, rows=rows, K=K, N=N, gy_stride=gy_stride, blocks=blocks)
    return [nw,
            RawOp(OP_REDUCE, {"dst": dst, "src": part},
                  dict(count=span, stride=span, splits=blocks, accumulate=0, ncols=0, act=0, beta=1.0), False,
                  name + ".fold", ReduceDesc)]


def split_rows(n_rt, want_ctas):
    """(splits, row tiles per split) with no empty split."""
    per = max(1, -(-n_rt // max(1, want_ctas)))
    return -(-n_rt // per), per


class CnnNet(object):
    """CNNBase (reference model.py:169-195: conv 8x8/4 -> 4x4/2 -> 3x3/1 -> fc, ReLU after each) and the GRU's
    input product on the TMA-fed engine.  Activations are NHWC planes (x, lo); the frames are re-laid out once per
    minibatch as a 21 x 21 x 64 space-to-depth tensor so that conv1 is a 2 x 2 stride-1 convolution."""

    def __init__(self, nb):
        self.nb = nb
        self.B, self.H = nb.B, nb.spec.hidden
        self.frames_lo = nb.obs_code != 1          # fp32 frames are not exact in TF32: they get a lo plane too
        # Training programs read minibatch rows of the rollout buffer through a runtime row list.  The frames do
        # not change during an update, so their space-to-depth re-layout is done ONCE per update for the whole
        # rollout (`prologue`) and conv1 gathers its images from it through the row list (TMA coordinates).
        self.full_rows = getattr(nb, "full_rows", 0) if nb.gather else 0
        self.prologue = []
        # the same re-layout restricted to the rows of ONE minibatch (row list, scattered to the rows' own slots):
        # recurrent updates run it before each minibatch of the first epoch, so that the frames of a minibatch's
        # environments can still be in flight from the host while earlier minibatches compute
        self.prologue_chunk = []

    # -- small helpers ------------------------------------------------------------------------------
    def buf2(self, name, numel):
        """x plane + lo plane."""
        return self.nb.buf(name, numel), self.nb.buf(name + "_lo", numel)

    def image(self, name, rows, cols):
        self.nb.buf(name, rows * cols)
        self.nb.buf(name + "_lo", rows * cols)
        return self.nb.n(name)

    def P(self, key):
        return ("P", self.nb.L.entries[key][0])

    def G(self, key):
        return ("G", self.nb.L.entries[key][0])

    # -- forward --------------------------------------------------------------------------------------
    def forward(self):
        nb, B, H = self.nb, self.B, self.H
        ops = []
        if self.full_rows:
            nb.sizes["x0_full"] = self.full_rows * 441 * 64
            x0 = ("x0_full", 0)
            x0_lo = None
            if self.frames_lo:
                nb.sizes["x0_full_lo"] = self.full_rows * 441 * 64
                x0_lo = ("x0_full_lo", 0)
        else:
            x0 = nb.buf("x0", B * 441 * 64)
            x0_lo = nb.buf("x0_lo", B * 441 * 64) if self.frames_lo else None
        # uint8 frames are exact in bf16: conv1's forward reads a bf16 copy of x0 (half the bytes, 64 channels per
        # 128-byte row) against a three-plane bf16 weight image; its weight gradient keeps the fp32 tensor
        self.c1_bf16 = not self.frames_lo
        x0b = None
        if self.c1_bf16:
            if self.full_rows:
                nb.sizes["x0b_full"] = self.full_rows * 441 * 32
                x0b = ("x0b_full", 0)
            else:
                x0b = nb.buf("x0b", B * 441 * 32)
        n_img = self.full_rows or B                  # images in the tensor conv1 reads
        gd = 3 if self.full_rows else 0              # ... gathered through the row list along the image dimension
        self.x0_imgs, self.x0_gather = n_img, gd
        (a1, a1_lo), (a2, a2_lo) = self.buf2("a1", B * 400 * 32), self.buf2("a2", B * 81 * 64)
        (a3, a3_lo), (h, h_lo) = self.buf2("a3", B * 49 * 32), self.buf2("h", B * H)
        self.refs = dict(x0=x0, x0_lo=x0_lo, a1=a1, a1_lo=a1_lo, a2=a2, a2_lo=a2_lo, a3=a3, a3_lo=a3_lo, h=h, h_lo=h_lo)

        # ---- weight images (weights change with every optimiser step) ----
        w1, w2, w3, wf = self.image("img_c1", 32, 256), self.image("img_c2", 64, 512), self.image("img_c3", 32, 576), \
            self.image("img_fc", H, 1568)
        # conv1: k = (by, bx | c, dy, dx) -> W[out, c, 4 by + dy, 4 bx + dx]
        k1 = [c * 64 + (4 * by + dy) * 8 + 4 * bx + dx for by in range(2) for bx in range(2) for c in range(4)
              for dy in range(4) for dx in range(4)]
        k2 = [c * 16 + ky * 4 + kx for ky in range(4) for kx in range(4) for c in range(32)]
        k3 = [c * 9 + ky * 3 + kx for ky in range(3) for kx in range(3) for c in range(64)]
        self.fc_perm = [c * 49 + y * 7 + x for y in range(7) for x in range(7) for c in range(32)]   # NHWC -> NCHW flatten
        if self.c1_bf16:
            w1b = self.image("img_c1b", 32, 256)
            nb.buf("img_c1b_3", 32 * 256)
            ops.append(prep_op("weight_image", w1b, self.P("base.main.0.weight"), 32, 256, [n * 256 for n in range(32)], k1,
                               bf16=True))
        ops += [prep_op("weight_image", w1, self.P("base.main.0.weight"), 32, 256, [n * 256 for n in range(32)], k1),
                prep_op("weight_image", w2, self.P("base.main.2.weight"), 64, 512, [n * 512 for n in range(64)], k2),
                prep_op("weight_image", w3, self.P("base.main.4.weight"), 32, 576, [n * 576 for n in range(32)], k3),
                prep_op("weight_image", wf, self.P("base.main.7.weight"), H, 1568, [n * 1568 for n in range(H)], self.fc_perm)]

        # ---- frames -> space-to-depth ----
        frames, is_u8, gather = (nb.obs, 0), nb.obs_code == 1, nb.gather

        ring = getattr(nb, "ring", None)
        ring_valid, ring_n = ring if ring is not None else (None, 0)
        assert ring is None or (self.full_rows and is_u8), "the frame ring feeds whole-rollout training programs"

        def make_s2d(arena):
            return S2dDesc(arena.ptr(x0), arena.ptr(x0_lo), arena.ptr(frames), None, 1 if is_u8 else 0, n_img, 4, 84, 84, 4,
                           arena.ptr(x0b), 0, ring_n, arena.ptr(ring_valid))
        s2d = TcxOp(OP_TCX_S2D, make_s2d, "frames_s2d", runtime_idx=gather and not self.full_rows)
        s2d.spec = dict(out=x0, out_lo=x0_lo, frames=frames, n_img=n_img)
        if self.full_rows:
            self.prologue.append(s2d)

            def make_s2d_chunk(arena):
                return S2dDesc(arena.ptr(x0), arena.ptr(x0_lo), arena.ptr(frames), None, 1 if is_u8 else 0, B, 4, 84, 84, 4,
                               arena.ptr(x0b), 1, ring_n, arena.ptr(ring_valid))
            chunk = TcxOp(OP_TCX_S2D, make_s2d_chunk, "frames_s2d_chunk", runtime_idx=True)
            chunk.spec = dict(out=x0, out_lo=x0_lo, frames=frames, n_img=B, scatter=True)
            self.prologue_chunk.append(chunk)
        else:
            ops.append(s2d)

        # ---- conv1: 2 x 2 taps x 64 channels, tile = 12 output rows of one image ----
        if self.c1_bf16:
            kbs = [((0, bx, by, 0), (by * 2 + bx) * 64) for by in range(2) for bx in range(2)]
            # patch mode: the 2 x 2 taps of a tile are windows of ONE 21 x 13 pixel patch (12 output rows + 1 halo row,
            # the full 21-pixel row pitch); accumulator row = patch pixel, the 21st column is junk (x >= o_ext)
            pm = os.environ.get("A2CB_TCX_PATCH", "1") != "0"
            ops.append(fwd_op("conv1.fwd", x0b, None, (64, 21, 21, n_img), (1, 64, 64 * 21, 64 * 441),
                              (64, 21, 13, 1) if pm else (64, 20, 12, 1), w1b, 32, 256,
                              kbs, tiling(2 * B, 2, 12, t_div=2, tq_dim=3, tq_step=1), a1, a1_lo, (0, 32, 640, 12800),
                              (0, 20, 20, B), 32, alpha=1.0 / 255.0, bias=self.P("base.main.0.bias"), act=1, chain=2,
                              flops=2.0 * B * 400 * 32 * 256, gather_dim=gd, bf16=True, emu_A=x0,
                              patch=(252, (64, 21, 12, 1)) if pm else None))
        else:
            kbs = [((c0, bx, by, 0), (by * 2 + bx) * 64 + c0) for by in range(2) for bx in range(2) for c0 in (0, 32)]
            ops.append(fwd_op("conv1.fwd", x0, x0_lo, (64, 21, 21, n_img), (1, 64, 64 * 21, 64 * 441), (32, 20, 12, 1), w1, 32, 256,
                              kbs, tiling(2 * B, 2, 12, t_div=2, tq_dim=3, tq_step=1), a1, a1_lo, (0, 32, 640, 12800),
                              (0, 20, 20, B), 32, alpha=1.0 / 255.0, bias=self.P("base.main.0.bias"), act=1,
                              flops=2.0 * B * 400 * 32 * 256, gather_dim=gd))
        # ---- conv2 (stride 2): parity view (px|c, X, py, Y, n) of a1, tile = 3 images ----
        # (2 images / 3 stages measured slower: these layers are bound by bytes moved into shared memory, and smaller
        # tiles re-load the weight tile more often)
        kbs = [(((kx % 2) * 32, kx // 2, ky % 2, ky // 2), (ky * 4 + kx) * 32) for ky in range(4) for kx in range(4)]
        k = 3
        ops.append(fwd_op("conv2.fwd", a1, a1_lo, (64, 10, 2, 10, B), (1, 64, 640, 1280, 12800), (32, 9, 1, 9, k), w2, 64, 512,
                          kbs, tiling(-(-B // k), 4, k), a2, a2_lo, (0, 64, 0, 576, 5184), (0, 9, 1, 9, B), 64,
                          bias=self.P("base.main.2.bias"), act=1, flops=2.0 * B * 81 * 64 * 512))
        # ---- conv3: tile = 5 images ----
        if os.environ.get("A2CB_TCX_PATCH3", "0") == "1":
            # patch mode (TF32, streamed weights): per tile and channel half ONE 9 x 9 x 3-image patch (x and lo planes);
            # the nine taps are row-shifted windows of it, accumulator row = patch pixel (x, y >= 7 are junk)
            kbs = [((c0, kx, ky, 0), (ky * 3 + kx) * 64 + c0) for c0 in (0, 32) for ky in range(3) for kx in range(3)]
            k = 3
            ops.append(fwd_op("conv3.fwd", a2, a2_lo, (64, 9, 9, B), (1, 64, 576, 5184), (32, 9, 9, k), w3, 32, 576, kbs,
                              tiling(-(-B // k), 3, k), a3, a3_lo, (0, 32, 224, 1568), (0, 7, 7, B), 32,
                              bias=self.P("base.main.4.bias"), act=1, chain=6, flops=2.0 * B * 49 * 32 * 576,
                              patch=(81 * k, (32, 9, 9, k))))
        else:
            kbs = [((c0, kx, ky, 0), (ky * 3 + kx) * 64 + c0) for ky in range(3) for kx in range(3) for c0 in (0, 32)]
            k = 5
            ops.append(fwd_op("conv3.fwd", a2, a2_lo, (64, 9, 9, B), (1, 64, 576, 5184), (32, 7, 7, k), w3, 32, 576, kbs,
                              tiling(-(-B // k), 3, k), a3, a3_lo, (0, 32, 224, 1568), (0, 7, 7, B), 32,
                              bias=self.P("base.main.4.bias"), act=1, chain=6, flops=2.0 * B * 49 * 32 * 576))
        # ---- fc ----
        ops.append(self.linear_fwd("fc.fwd", a3, a3_lo, 1568, wf, H, h, h_lo, bias=self.P("base.main.7.bias"), act=1))
        return ops, h

    def linear_fwd(self, name, x, x_lo, K, img, N, out, out_lo, bias=None, act=0, mask=None):
        B = self.B
        R = 256 if B > 4096 else 128
        kbs = [((32 * i, 0, 0, 0), 32 * i) for i in range(K // 32)]
        return fwd_op(name, x, x_lo, (K, B), (1, K), (32, R), img, N, K, kbs, tiling(-(-B // R), 1, R), out, out_lo,
                      (0, N), (0, B), N, bias=bias, act=act, mask=mask, chain=8, flops=2.0 * B * N * K)

    def gru_ih_forward(self, gi):
        """gi = h W_ih^T + b_ih for all rows (model.py:113 / torch GRU input product)."""
        H = self.H
        wi = self.image("img_gih", 3 * H, H)
        r = self.refs
        return [prep_op("weight_image", wi, self.P("base.gru.weight_ih_l0"), 3 * H, H, [n * H for n in range(3 * H)],
                        list(range(H))),
                self.linear_fwd("gru_ih.fwd", r["h"], r["h_lo"], H, wi, 3 * H, gi, None, bias=self.P("base.gru.bias_ih_l0"))]

    # -- backward -------------------------------------------------------------------------------------
    def bias_grad(self, name, x, rows, C, dst, lo=None):
        """Bias gradient = column sums of the output gradient (lo: also emit its lo plane in the same pass)."""
        nb = self.nb
        blocks = colsum_blocks(rows)
        part = nb.buf("bpart_" + name, blocks * C)
        from ._engine import RawOp, ReduceDesc, OP_REDUCE
        return [colsum_op(name + ".bias_colsum", part, x, rows, C, C, lo=lo),
                RawOp(OP_REDUCE, {"dst": dst, "src": part},
                      dict(count=C, stride=C, splits=blocks, accumulate=0, ncols=0, act=0, beta=1.0), False,
                      name + ".bias_fold", ReduceDesc)]

    def fold(self, name, part, dst, count, splits):
        from ._engine import RawOp, ReduceDesc, OP_REDUCE
        return RawOp(OP_REDUCE, {"dst": dst, "src": part},
                     dict(count=count, stride=count, splits=splits, accumulate=0, ncols=0, act=0, beta=1.0), False,
                     name + ".fold", ReduceDesc)

    def linear_wgrad(self, name, x, x_lo, K, gy, gy_lo, N, dst, perm=None):
        """dW[n][perm(k)] = sum_rows gy[row, n] x[row, k]."""
        B = self.B
        R = 32
        ot_a, ot_b = -(-K // 128), -(-N // 256)
        n_rt = -(-B // R)
        splits, per = split_rows(n_rt, max(1, 148 // (ot_a * ot_b)))      # one wave of CTAs (fc: 26 tiles x 5 splits, not 6)
        part = self.nb.buf("wpart_" + name, splits * N * K)
        row_off = []
        for f in range(ot_a * 128):
            row_off.append(-1 if f >= K else (perm[f] if perm is not None else f))
        groups = [[(32 * j, 0, 0, 0) for j in range(4)]]
        b_boxes = [(32 * j, 0, 0, 0) for j in range(8)]
        op = wgrad_op(name, x, x_lo, (K, B), (1, K), (32, R), gy, gy_lo, (N, B), (1, N), (32, R), groups, b_boxes,
                      tiling(n_rt, 1, R), splits, 16, ot_a, ot_b, 128, 256, part, N * K, row_off, K, N,
                      flops=2.0 * B * N * K)
        return [op, self.fold(name, part, dst, N * K, splits)]

    def backward(self, gh, gh_lo):
        """gh: gradient w.r.t. the trunk output h (already masked by relu'), both planes."""
        nb, B, H, r = self.nb, self.B, self.H, self.refs
        ops = []
        (g3, g3_lo), (g2, g2_lo), (g1, g1_lo) = self.buf2("g3", B * 49 * 32), self.buf2("g2", B * 81 * 64), \
            self.buf2("g1", B * 400 * 32)
        # ---- data-gradient weight images ----
        wfd, w3d, w2d = self.image("img_fc_d", 1568, H), self.image("img_c3_d", 64, 288), self.image("img_c2_d", 128, 256)
        ops += [prep_op("weight_image", wfd, self.P("base.main.7.weight"), 1568, H, self.fc_perm, [k * 1568 for k in range(H)]),
                # [ci][(ky, kx | co)] = W3[co, ci, ky, kx]
                prep_op("weight_image", w3d, self.P("base.main.4.weight"), 64, 288, [ci * 9 for ci in range(64)],
                        [co * 576 + ky * 3 + kx for ky in range(3) for kx in range(3) for co in range(32)]),
                # [(py, px | ci)][(b, a | co)] = W2[co, ci, py + 2 b, px + 2 a]
                prep_op("weight_image", w2d, self.P("base.main.2.weight"), 128, 256,
                        [ci * 16 + py * 4 + px for py in range(2) for px in range(2) for ci in range(32)],
                        [co * 512 + (2 * b) * 4 + 2 * a for b in range(2) for a in range(2) for co in range(64)])]
        # ---- fc ----
        ops += self.linear_wgrad("fc.wgrad", r["a3"], r["a3_lo"], 1568, gh, gh_lo, H, self.G("base.main.7.weight"),
                                 perm=self.fc_perm)
        ops += self.bias_grad("fc", gh, B, H, self.G("base.main.7.bias"))
        ops.append(self.linear_fwd("fc.dgrad", gh, gh_lo, H, wfd, 1568, g3, g3_lo, mask=r["a3"]))
        # ---- conv3 ----
        ops += self.conv3_wgrad(g3, g3_lo)
        ops += self.bias_grad("conv3", g3, B * 49, 32, self.G("base.main.4.bias"))
        kbs = [((0, -kx, -ky, 0), (ky * 3 + kx) * 32) for ky in range(3) for kx in range(3)]
        k = 3
        # conv2's bias gradient = column sums of g2: accumulated by the epilogue that writes g2 (one partial row per CTA)
        n_t = -(-B // k)
        bp2 = nb.buf("bpart_conv2", fwd_grid(n_t, 64) * 64)
        ops.append(fwd_op("conv3.dgrad", g3, g3_lo, (32, 7, 7, B), (1, 32, 224, 1568), (32, 9, 9, k), w3d, 64, 288, kbs,
                          tiling(n_t, 3, k), g2, g2_lo, (0, 64, 576, 5184), (0, 9, 9, B), 64, mask=r["a2"], chain=5,
                          flops=2.0 * B * 49 * 32 * 576, colsum_part=bp2))
        ops.append(self.fold("conv2.bias", bp2, self.G("base.main.2.bias"), 64, fwd_grid(n_t, 64)))
        # ---- conv2 ----
        ops += self.conv2_wgrad(g2, g2_lo)
        # the four stride-parity classes (py, px) of the transposed convolution read the SAME boxes of g2 (tap (b, a)
        # -> rows (u - b, v - a)); only the weights differ, so they are 4 x 32 columns of one product
        k = 1          # 100 rows / tile: one M block, 64 KiB stages -> a 3-deep pipeline (2 images would leave only 2 stages)
        kbs = [((c0, -a, -b, 0), (b * 2 + a) * 64 + c0) for b in range(2) for a in range(2) for c0 in (0, 32)]
        n_t = -(-B // k)
        bp1 = nb.buf("bpart_conv1", fwd_grid(n_t, 128) * 128)
        ops.append(fwd_op("conv2.dgrad", g2, g2_lo, (64, 9, 9, B), (1, 64, 576, 5184), (32, 10, 10, k), w2d, 128, 256,
                          kbs, tiling(n_t, 3, k), g1, g1_lo, (0, 64, 1280, 12800), (0, 10, 10, B), 128,
                          mask=r["a1"], flops=2.0 * B * 81 * 64 * 512,
                          col_blocks=(32, [(py * 20 + px) * 32 for py in range(2) for px in range(2)]), colsum_part=bp1))
        # conv1's bias gradient: the partial rows hold 4 parity classes x 32 channels -> fold over CTAs and classes
        ops.append(self.fold("conv1.bias", bp1, self.G("base.main.0.bias"), 32, 4 * fwd_grid(n_t, 128)))
        # ---- conv1 ----
        ops += self.conv1_wgrad(g1, g1_lo)
        return ops

    def conv3_wgrad(self, g3, g3_lo):
        B, r = self.B, self.refs
        groups, row_off = [], []
        for g in range(5):
            boxes = []
            for j in range(4):
                tap = 2 * g + j // 2
                boxes.append(((j % 2) * 32, tap % 3, tap // 3, 0) if tap < 9 else (0, 0, 0, OOB))
                row_off += [(((j % 2) * 32 + c) * 9 + tap) if tap < 9 else -1 for c in range(32)]
            groups.append(boxes)
        splits, per = split_rows(B, 148)
        part = self.nb.buf("wpart_conv3", splits * 32 * 576)
        op = wgrad_op("conv3.wgrad", r["a2"], r["a2_lo"], (64, 9, 9, B), (1, 64, 576, 5184), (32, 7, 7, 1), g3, g3_lo,
                      (32, 7, 7, B), (1, 32, 224, 1568), (32, 7, 7, 1), groups, [(0, 0, 0, 0)], tiling(B, 3, 1), splits, 8,
                      1, 1, 0, 0, part, 32 * 576, row_off, 576, 32, flops=2.0 * B * 49 * 32 * 576)
        return [op, self.fold("conv3.wgrad", part, self.G("base.main.4.weight"), 32 * 576, splits)]

    def conv2_wgrad(self, g2, g2_lo):
        B, r = self.B, self.refs
        groups, row_off = [], []
        for ky in range(4):
            boxes = []
            for kx in range(4):
                boxes.append(((kx % 2) * 32, kx // 2, ky % 2, ky // 2))
                row_off += [c * 16 + ky * 4 + kx for c in range(32)]
            groups.append(boxes)
        n_rt = 2 * B                                   # half images: 5 output rows x 9
        splits, per = split_rows(n_rt, 148)
        part = self.nb.buf("wpart_conv2", splits * 64 * 512)
        op = wgrad_op("conv2.wgrad", r["a1"], r["a1_lo"], (64, 10, 2, 10, B), (1, 64, 640, 1280, 12800), (32, 9, 1, 5, 1),
                      g2, g2_lo, (64, 9, 1, 9, B), (1, 64, 64, 576, 5184), (32, 9, 1, 5, 1), groups,
                      [(0, 0, 0, 0), (32, 0, 0, 0)], tiling(n_rt, 3, 5, t_div=2, tq_dim=4, tq_step=1), splits, 8, 1, 1, 0, 0,
                      part, 64 * 512, row_off, 512, 64, flops=2.0 * B * 81 * 64 * 512)
        return [op, self.fold("conv2.wgrad", part, self.G("base.main.2.weight"), 64 * 512, splits)]

    def conv1_wgrad(self, g1, g1_lo):
        B, r = self.B, self.refs
        groups, row_off = [], []
        for g in range(2):
            boxes = []
            for j in range(4):
                by, bx = g, j // 2
                boxes.append(((j % 2) * 32, bx, by, 0))
                for cl in range(32):
                    cc = (j % 2) * 32 + cl
                    c, dy, dx = cc // 16, (cc // 4) % 4, cc % 4
                    row_off.append(c * 64 + (4 * by + dy) * 8 + 4 * bx + dx)
            groups.append(boxes)
        yr = 2 if self.frames_lo else 4                 # output rows per row tile
        n_rt = (20 // yr) * B
        splits, per = split_rows(n_rt, 148)
        part = self.nb.buf("wpart_conv1", splits * 32 * 256)
        op = wgrad_op("conv1.wgrad", r["x0"], r["x0_lo"], (64, 21, 21, self.x0_imgs), (1, 64, 64 * 21, 64 * 441), (32, 20, yr, 1),
                      g1, g1_lo, (32, 20, 20, B), (1, 32, 640, 12800), (32, 20, yr, 1), groups, [(0, 0, 0, 0)],
                      tiling(n_rt, 2, yr, t_div=20 // yr, tq_dim=3, tq_step=1), splits, 6 if yr == 4 else 12, 1, 1, 0, 0, part,
                      32 * 256, row_off, 256, 32, alpha=1.0 / 255.0, flops=2.0 * B * 400 * 32 * 256,
                      gather_dim=self.x0_gather)
        return [op, self.fold("conv1.wgrad", part, self.G("base.main.0.weight"), 32 * 256, splits)]

    # -- GRU linear parts (the recurrence itself is csrc/gru_persistent.cu) ---------------------------
    def gru_backward(self, dgi, dgh, hm, gx, gx_lo):
        """Weight / bias gradients of both GRU matrices and the gradient w.r.t. the GRU input (masked by relu' of h)."""
        nb, B, H, r = self.nb, self.B, self.H, self.refs
        dgi_lo, dgh_lo, hm_lo = nb.buf("gru_dgi_lo", B * 3 * H), nb.buf("gru_dgh_lo", B * 3 * H), nb.buf("gru_hm_lo", B * H)
        wid = self.image("img_gih_d", H, 3 * H)
        # bias gradients (column sums) and the lo planes of the gate gradients come out of one pass over them
        ops = [lo_op("lo_plane", hm, hm_lo, B * H)]
        ops += self.bias_grad("gru_hh", dgh, B, 3 * H, self.G("base.gru.bias_hh_l0"), lo=dgh_lo)
        ops += self.bias_grad("gru_ih", dgi, B, 3 * H, self.G("base.gru.bias_ih_l0"), lo=dgi_lo)
        ops += self.linear_wgrad("gru_hh.wgrad", hm, hm_lo, H, dgh, dgh_lo, 3 * H, self.G("base.gru.weight_hh_l0"))
        ops += self.linear_wgrad("gru_ih.wgrad", r["h"], r["h_lo"], H, dgi, dgi_lo, 3 * H, self.G("base.gru.weight_ih_l0"))
        ops.append(prep_op("weight_image", wid, self.P("base.gru.weight_ih_l0"), H, 3 * H, list(range(H)),
                           [k * H for k in range(3 * H)]))
        ops.append(self.linear_fwd("gru_ih.dgrad", dgi, dgi_lo, 3 * H, wid, H, gx, gx_lo, mask=r["h"]))
        return ops


# ---------------------------------------------------------------------------------------------------
# numpy emulator of the ops above (CPU tests of the host-side geometry: boxes, tap tables, tilings, weight
# images, row-offset tables).  Planes: the x plane is treated as the exact fp32 value, lo planes are ignored.
# ---------------------------------------------------------------------------------------------------
def _np_view(bufs, ref):
    name, off = ref
    return bufs[name][off:]


def _tile_base(t, tile):
    base = [0, 0, 0, 0, 0]
    tq, tr = divmod(tile, t.t_div)
    base[t.tr_dim] += tr * t.tr_step
    base[t.tq_dim] += tq * t.tq_step
    return base


def _box_rows(box):
    """Box-local indices (i1..i4) of every row, dimension 1 fastest."""
    import numpy as np
    r = np.arange(box[1] * box[2] * box[3] * box[4])
    i1 = r % box[1]; r = r // box[1]
    i2 = r % box[2]; r = r // box[2]
    i3 = r % box[3]; i4 = r // box[3]
    return i1, i2, i3, i4


def _load_box(src, dims, strides, box, c, fill=0.0, cw=32):
    """[rows, cw] values of the box whose corner has coordinates c (5 ints); out-of-bounds elements are zero."""
    import numpy as np
    ii = _box_rows(box)
    ok = np.ones(ii[0].shape, bool)
    off = np.zeros(ii[0].shape, np.int64)
    for d in range(1, 5):
        g = c[d] + ii[d - 1]
        ok &= (g >= 0) & (g < dims[d])
        off += np.clip(g, 0, dims[d] - 1) * strides[d]
    ch = c[0] + np.arange(cw)
    okc = (ch >= 0) & (ch < dims[0])
    vals = src[off[:, None] + np.clip(ch, 0, dims[0] - 1)[None, :]].astype(np.float64)
    return np.where(ok[:, None] & okc[None, :], vals, fill)


def _pad5(dims, strides, box):
    dims, strides, box = list(dims), list(strides), list(box)
    while len(dims) < 5:
        strides.append(strides[-1] * dims[-1]); dims.append(1); box.append(1)
    return dims, strides, box


def emulate(op, bufs, idx=None):
    """Executes one plan-list entry on numpy buffers (dict name -> flat array).  idx: runtime row list."""
    import numpy as np
    from . import _engine
    if isinstance(op, _engine.RawOp):
        assert op.kind == _engine.OP_REDUCE, op.kind
        dst, src = _np_view(bufs, op.fields["dst"]), _np_view(bufs, op.fields["src"])
        cnt, stride, splits = op.ints["count"], op.ints["stride"], op.ints["splits"]
        dst[:cnt] = sum(src[s * stride:s * stride + cnt].astype(np.float64) for s in range(splits)).astype(np.float32)
        return
    k, sp = op.kind, getattr(op, "spec", None)
    if k == OP_TCX_PREP:
        img, src, N, K, ntab, ktab = op.prep
        nt, kt = np.asarray(ntab)[:, None], np.asarray(ktab)[None, :]
        s = _np_view(bufs, src)
        bufs[img][:N * K] = np.where((nt >= 0) & (kt >= 0), s[np.maximum(nt, 0) + np.maximum(kt, 0)], 0.0).ravel()
    elif k == OP_TCX_S2D:
        fr = _np_view(bufs, sp["frames"])
        n = sp["n_img"]
        rows = np.asarray(idx) if (op.runtime_idx and idx is not None) else np.arange(n)
        f = fr.reshape(-1, 4, 21, 4, 21, 4)[rows].astype(np.float32)            # [b, c, by, dy, bx, dx]
        f = f.transpose(0, 2, 4, 1, 3, 5).reshape(len(rows), 28224)
        if sp.get("scatter"):
            _np_view(bufs, sp["out"]).reshape(-1, 28224)[rows] = f              # image b -> slot rows[b]
        else:
            _np_view(bufs, sp["out"])[:n * 28224] = f.reshape(-1)
    elif k == OP_TCX_LO:
        pass
    elif k == OP_TCX_COLSUM:
        x = _np_view(bufs, sp["x"])
        rows, C, pitch = sp["rows"], sp["C"], sp["pitch"]
        m = np.lib.stride_tricks.as_strided(x, (rows, C), (pitch * x.itemsize, x.itemsize))
        part = _np_view(bufs, sp["part"])
        part[:sp["blocks"] * C] = 0
        part[:C] = m.astype(np.float64).sum(0)
    elif k == OP_TCX_NARROW_WGRAD:
        x, gy = _np_view(bufs, sp["x"]), _np_view(bufs, sp["gy"])
        rows, K, N, gs = sp["rows"], sp["K"], sp["N"], sp["gy_stride"]
        X = x[:rows * K].reshape(rows, K).astype(np.float64)
        G = np.lib.stride_tricks.as_strided(gy, (rows, N), (gs * gy.itemsize, gy.itemsize)).astype(np.float64)
        part = _np_view(bufs, sp["part"])
        part[:sp["blocks"] * (N * K + N)] = 0
        part[:N * K] = (G.T @ X).ravel()
        part[N * K:N * K + N] = G.sum(0)
    elif k == OP_TCX_FWD:
        dims, strides, box = _pad5(sp["dims"], sp["strides"], sp["box"])
        A = _np_view(bufs, sp["A"])
        img = bufs[sp["img"]][:sp["img_rows"] * sp["img_cols"]].reshape(sp["img_rows"], sp["img_cols"]).astype(np.float64)
        out = _np_view(bufs, sp["out"])
        mask = _np_view(bufs, sp["mask"]) if sp["mask"] is not None else None
        bias = _np_view(bufs, sp["bias"])[:sp["N"]].astype(np.float64) if sp["bias"] is not None else 0.0
        os_ = list(sp["o_stride"]) + [0] * (5 - len(sp["o_stride"]))
        oe = list(sp["o_ext"]) + [1] * (5 - len(sp["o_ext"]))
        ii = _box_rows(box)
        N = sp["N"]
        csum = np.zeros(N)
        for tile in range(sp["tiles"].n_tiles):
            base = _tile_base(sp["tiles"], tile)
            ba = list(base)
            if sp["gather_dim"]:
                ba[sp["gather_dim"]] = int(idx[base[sp["gather_dim"]]])
            acc = np.zeros((len(ii[0]), N))
            for a, b in sp["kbs"]:
                c = [a[0], a[1] + ba[1], a[2] + ba[2], a[3] + ba[3], ba[4]]
                acc += _load_box(A, dims, strides, box, c, cw=sp.get("cw", 32)) @ img[:N, b:b + sp.get("cw", 32)].T
            v = acc * sp["alpha"] + bias
            if sp["act"] == 1:
                v = np.maximum(v, 0.0)
            g = [base[d + 1] + ii[d] for d in range(4)]
            ok = np.ones(len(ii[0]), bool)
            off = np.full(len(ii[0]), sp["o_base"], np.int64)
            for d in range(4):
                ok &= g[d] < oe[d + 1]
                off += g[d] * os_[d + 1]
            cb = sp["col_blocks"]
            for n in range(N):
                o = off + (cb[1][n // cb[0]] + n % cb[0] if cb else n)
                col = v[:, n]
                if mask is not None:
                    col = np.where(mask[np.where(ok, o, 0)] > 0, col, 0.0)
                out[o[ok]] = col[ok].astype(np.float32)
                if sp.get("colsum_part") is not None:
                    csum[n] += col[ok].sum()
        if sp.get("colsum_part") is not None:
            part = _np_view(bufs, sp["colsum_part"])
            part[:fwd_grid(sp["tiles"].n_tiles, N) * N] = 0
            part[:N] = csum
    elif k == OP_TCX_WGRAD:
        ad, as_, ab = _pad5(sp["a_dims"], sp["a_strides"], sp["a_box"])
        bd, bs, bb = _pad5(sp["b_dims"], sp["b_strides"], sp["b_box"])
        A, Bm = _np_view(bufs, sp["A"]), _np_view(bufs, sp["Bm"])
        part = _np_view(bufs, sp["partial"])
        part[:sp["splits"] * sp["split_stride"]] = 0
        G, nbb = len(sp["groups"]), len(sp["b_boxes"])
        row_off = np.asarray(sp["row_off"])
        for ota in range(sp["ot_a"]):
            for otb in range(sp["ot_b"]):
                D = np.zeros((G * 128, 32 * nbb))
                for rt in range(sp["rows"].n_tiles):
                    base = _tile_base(sp["rows"], rt)
                    ba = list(base)
                    if sp["gather_dim"]:
                        ba[sp["gather_dim"]] = int(idx[base[sp["gather_dim"]]])
                    Bt = np.concatenate([_load_box(Bm, bd, bs, bb, [c[0] + otb * sp["ot_b_step"], c[1] + base[1], c[2] + base[2],
                                                                      c[3] + base[3], base[4]]) for c in sp["b_boxes"]], 1)
                    for g, boxes in enumerate(sp["groups"]):
                        At = np.concatenate([_load_box(A, ad, as_, ab, [c[0] + ota * sp["ot_a_step"], c[1] + ba[1], c[2] + ba[2],
                                                                         c[3] + ba[3], ba[4]]) for c in boxes], 1)
                        D[g * 128:(g + 1) * 128] += At.T @ Bt
                ro = row_off[ota * G * 128:(ota + 1) * G * 128]
                for n in range(32 * nbb):
                    col = otb * 32 * nbb + n
                    if col < sp["n_cols"]:
                        part[ro[ro >= 0] + col * sp["col_stride"]] = (D[ro >= 0, n] * sp["alpha"]).astype(np.float32)
    else:
        raise NotImplementedError(k)
