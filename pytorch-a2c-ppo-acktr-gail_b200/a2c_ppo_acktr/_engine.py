"""Execution engine: turns a policy + batch shape into op lists for liba2cb.so and runs them.

Host-side only (no tensor math): owns the activation / gradient arenas, resolves the
symbolic plans of `_plans.py` into `a2cb_gemm_t` descriptors holding device pointers,
and assembles the per-minibatch programs

    forward GEMMs -> fused loss epilogue -> backward GEMMs -> split-K folds -> clip + optimiser

that `a2cb_run_ops` launches from C in one call.  torch is used for device memory only.
"""
import ctypes
import math
import os

import torch

from . import _native
from . import _plans as P

c_i32, c_i64, c_f32, c_vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p

OP_GEMM, OP_REDUCE, OP_LOSS, OP_GRAD_NORM, OP_OPTIM, OP_FILL = 1, 2, 3, 4, 5, 6
OP_GRU_INIT, OP_GRU_FWD, OP_GRU_BWD, OP_SPATIAL_SUM, OP_GEMM_PREP_B = 7, 8, 9, 10, 11
OP_GRU_SEQ_FWD, OP_GRU_SEQ_BWD = 12, 13
RUNTIME_IDX = 1
SENTINEL = 1          # non-NULL placeholder for "use the runtime minibatch index list"


class LossDesc(ctypes.Structure):
    """Mirror of a2cb_loss_t."""
    _fields_ = [
        ("z", c_vp), ("logstd", c_vp), ("idx", c_vp), ("actions", c_vp), ("old_logp", c_vp),
        ("old_value", c_vp), ("returns", c_vp), ("adv", c_vp), ("noise", c_vp), ("dz", c_vp),
        ("dlogstd", c_vp), ("logp_out", c_vp), ("ent_out", c_vp), ("loss_out", c_vp), ("loss_accum", c_vp),
        ("scratch", c_vp),
        ("B", c_i32), ("A", c_i32), ("z_stride", c_i32), ("dz_stride", c_i32), ("act_stride", c_i32),
        ("dlogstd_stride", c_i32), ("dist", c_i32), ("mode", c_i32), ("use_clipped_value_loss", c_i32),
        ("clip", c_f32), ("c_v", c_f32), ("c_e", c_f32), ("inv_B", c_f32), ("n_valid", c_i32), ("valid_mod", c_i32), ("pad2_", c_i32),
        ("dlogstd_rows", c_vp),
    ]


class ReduceDesc(ctypes.Structure):
    _fields_ = [("dst", c_vp), ("src", c_vp), ("count", c_i64), ("stride", c_i64), ("splits", c_i32),
                ("accumulate", c_i32), ("bias", c_vp), ("ncols", c_i32), ("act", c_i32), ("beta", c_f32),
                ("pad_", c_i32)]


class GradNormDesc(ctypes.Structure):
    _fields_ = [("grads", c_vp), ("n", c_i64), ("scratch", c_vp), ("norm_out", c_vp)]


class OptimDesc(ctypes.Structure):
    _fields_ = [("params", c_vp), ("grads", c_vp), ("state1", c_vp), ("state2", c_vp), ("n", c_i64),
                ("norm", c_vp), ("kind", c_i32), ("first_step", c_i32),
                ("lr_eff", c_f32), ("eps", c_f32), ("b1", c_f32), ("b2", c_f32), ("omb1", c_f32), ("omb2", c_f32),
                ("bc2_sqrt", c_f32), ("max_norm", c_f32), ("dyn", c_vp)]


class FillDesc(ctypes.Structure):
    _fields_ = [("dst", c_vp), ("bytes", c_i64), ("value", c_i32), ("pad_", c_i32)]


class GruDesc(ctypes.Structure):
    """Mirror of a2cb_gru_t."""
    _fields_ = [("gi", c_vp), ("gh", c_vp), ("hm", c_vp), ("h0", c_vp), ("masks", c_vp), ("idx", c_vp),
                ("out", c_vp), ("r", c_vp), ("z", c_vp), ("n", c_vp), ("dout", c_vp), ("dgi", c_vp), ("dgh", c_vp),
                ("dcarry", c_vp), ("gh_part", c_vp), ("gh_bias", c_vp), ("dc_part", c_vp),
                ("T", c_i32), ("N", c_i32), ("H", c_i32), ("t", c_i32), ("gh_splits", c_i32), ("dc_splits", c_i32)]


class SpatialSumDesc(ctypes.Structure):
    _fields_ = [("out", c_vp), ("in_", c_vp), ("img_stride", c_i64), ("pitch", c_i64), ("B", c_i32), ("P", c_i32),
                ("C", c_i32), ("ow", c_i32)]


class MlpDesc(ctypes.Structure):
    """Mirror of a2cb_mlp_t."""
    _fields_ = [("P", c_vp), ("G", c_vp), ("obs", c_vp), ("idx", c_vp), ("acts", c_vp), ("z", c_vp), ("dz", c_vp),
                ("partial", c_vp)] + [(k, c_i64) for k in ("off_a0w", "off_a0b", "off_a2w", "off_a2b", "off_c0w", "off_c0b",
                                                            "off_c2w", "off_c2b", "off_hw", "off_hb", "span")] + \
               [("B", c_i32), ("d_in", c_i32), ("H", c_i32), ("zs", c_i32)]


OP_MLP_FWD, OP_MLP_BWD = 22, 23
OP_SAMPLE = 26


class SampleDesc(ctypes.Structure):
    """Mirror of a2cb_sample_t."""
    _fields_ = [("z", c_vp), ("logstd", c_vp), ("value_out", c_vp), ("action_out", c_vp), ("logp_out", c_vp),
                ("rng", c_vp), ("B", c_i32), ("A", c_i32), ("z_stride", c_i32), ("dist", c_i32),
                ("deterministic", c_i32), ("pad_", c_i32)]


def _pointer_sites(obj, lo, hi, out):
    """Every c_void_p field of a descriptor (nested structures and arrays of structures included) whose value
    lies in [lo, hi): appended to `out` as (owner, field name, offset from lo)."""
    for name, typ in obj._fields_:
        if typ is c_vp:
            v = getattr(obj, name)
            if v is not None and lo <= v < hi:
                out.append((obj, name, v - lo))
        elif isinstance(typ, type) and issubclass(typ, ctypes.Structure):
            _pointer_sites(getattr(obj, name), lo, hi, out)
        elif isinstance(typ, type) and issubclass(typ, ctypes.Array) and issubclass(typ._type_, ctypes.Structure):
            for el in getattr(obj, name):
                _pointer_sites(el, lo, hi, out)


class ActPlan(object):
    """No-grad forward + sampling head of one actor step for a fixed batch size (Policy.act / Policy.act_into).
    The program is built once against arena-owned placeholder tensors; `sites[role]` lists every descriptor field
    that points into a placeholder, so a call can re-point inputs and outputs (e.g. at the [step] slots of the
    rollout buffer) by patching a handful of pointers -- no staging copies, no per-step program."""

    ROLES_IN = ("obs", "hxs", "masks")
    ROLES_OUT = ("value", "action", "logp", "hout")

    def __init__(self, prog, sample, tensors, sites):
        self.prog, self.sample, self.tensors, self.sites = prog, sample, tensors, sites
        self.home = {r: t.data_ptr() for r, t in tensors.items()}
        self.current = dict(self.home)
        self.graphs = {}            # (mode, pointers) -> "seen" | (CUDA graph, launches): Policy._act_graphed

    def point(self, role, address):
        if self.current[role] != address:
            for owner, field, off in self.sites[role]:
                setattr(owner, field, address + off)
            self.current[role] = address

    def run(self, deterministic, pointers=None):
        for role, base in self.home.items():
            self.point(role, base if pointers is None else pointers.get(role, base))
        self.sample.deterministic = int(bool(deterministic))
        self.prog.run()


class GruSeqDesc(ctypes.Structure):
    """Mirror of a2cb_gru_seq_t: the per-step descriptor + the recurrent weights."""
    _fields_ = GruDesc._fields_ + [("W_hh", c_vp), ("b_hh", c_vp)]


def gru_seq_supported(N, H):
    """Persistent whole-sequence GRU kernel usable?  (A2CB_GRU=steps forces the per-step path.)"""
    import os
    if os.environ.get("A2CB_GRU", "") == "steps":
        return False
    return bool(_native.lib().a2cb_gru_seq_supported(int(N), int(H)))


class RawOp(object):
    """A non-GEMM op in a plan list: kind + symbolic fields resolved against the arena."""

    def __init__(self, kind, fields, ints, runtime_idx=False, name="op", struct=None):
        self.kind, self.fields, self.ints, self.runtime_idx, self.name = kind, fields, ints, runtime_idx, name
        self.struct = struct


class Op(ctypes.Structure):
    _fields_ = [("kind", c_i32), ("flags", c_i32), ("desc", c_vp)]


_native.register({
    "a2cb_gemm": (c_i32, [ctypes.POINTER(P.GemmDesc), c_vp]),
    "a2cb_gemm_prep_b": (c_i32, [ctypes.POINTER(P.GemmDesc), c_vp]),
    "a2cb_gemm_b_image_floats": (c_i64, [ctypes.POINTER(P.GemmDesc)]),
    "a2cb_reduce_splits": (c_i32, [c_vp, c_vp, c_i64, c_i64, c_i32, c_i32, c_vp, c_i32, c_i32, c_f32, c_vp]),
    "a2cb_spatial_sum": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i64, c_i64, c_vp]),
    "a2cb_kfac_scale": (c_i32, [c_vp, c_vp, c_vp, c_i32, c_i32, c_f32, c_vp]),
    "a2cb_kfac_nu": (c_i32, [c_vp, c_vp, c_i64, c_vp, c_vp, c_f32, c_f32, c_vp]),
    "a2cb_frames_to_f32": (c_i32, [c_vp, c_vp, c_i64, c_i32, c_f32, c_vp]),
    "a2cb_loss_epilogue": (c_i32, [ctypes.POINTER(LossDesc), c_vp]),
    "a2cb_sample_actions": (c_i32, [ctypes.POINTER(SampleDesc), c_vp]),
    "a2cb_loss_scratch_floats": (c_i64, [c_i32]),
    "a2cb_adv_moments": (c_i32, [c_vp, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "a2cb_adv_normalize": (c_i32, [c_vp, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "a2cb_grad_norm": (c_i32, [c_vp, c_i64, c_vp, c_vp, c_vp]),
    "a2cb_optim_step": (c_i32, [c_i32, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp] + [c_f32] * 8 + [c_i32, c_vp]),
    "a2cb_run_ops": (c_i32, [ctypes.POINTER(Op), c_i32, c_vp, c_vp]),
    "a2cb_gru_op": (c_i32, [c_i32, ctypes.POINTER(GruDesc), c_vp]),
    "a2cb_gru_seq": (c_i32, [c_i32, c_vp, c_vp]),
    "a2cb_mlp_fused": (c_i32, [c_i32, ctypes.POINTER(MlpDesc), c_vp]),
    "a2cb_mlp_fused_supported": (c_i32, [c_i32, c_i32, c_i32]),
    "a2cb_gru_seq_supported": (c_i32, [c_i32, c_i32]),
})


def _aff(t):
    return P.Affine3(*t)


class Arena(object):
    """Named flat device buffers."""

    def __init__(self, device):
        self.device = device
        self.bufs = {}
        self._consts = {}           # (dtype, bytes) -> device tensor, see const()

    def bind(self, name, tensor):
        self.bufs[name] = tensor

    def const(self, host_array):
        """Device copy of a small read-only host table (numpy array), shared by content: programs keep raw device
        pointers to these tables, so a table must stay alive -- and unmoved -- for as long as ANY program built
        from it does; rebuilding a program therefore never replaces a table."""
        key = (host_array.dtype.str, host_array.tobytes())
        t = self._consts.get(key)
        if t is None:
            t = torch.from_numpy(host_array.copy()).to(self.device)
            self._consts[key] = t
        return t

    def stage(self, name, src):
        """Copies `src` into an arena-owned buffer of the same size and dtype (one per name) and returns it: the
        programs that read it are then independent of where the caller's tensor lives."""
        t = self.bufs.get(name)
        if t is None or t.numel() != src.numel() or t.dtype != src.dtype:
            t = torch.empty(src.numel(), dtype=src.dtype, device=self.device)
            self.bufs[name] = t
        t.copy_(src.reshape(-1))
        return t

    def ensure(self, name, numel, dtype=torch.float32):
        t = self.bufs.get(name)
        if t is None or t.numel() < numel or t.dtype != dtype:
            t = torch.zeros(max(int(numel), 1), dtype=dtype, device=self.device)
            self.bufs[name] = t
        return t

    def ptr(self, ref):
        """(name, element offset) -> device address."""
        if ref is None:
            return None
        name, off = ref
        t = self.bufs[name]
        assert 0 <= off <= t.numel(), (name, off, t.numel())
        return t.data_ptr() + off * t.element_size()


def resolve(plan, arena):
    """Plan (symbolic) -> a2cb_gemm_t with device pointers."""
    g = P.GemmDesc()
    g.A, g.B, g.C = arena.ptr(plan.A), arena.ptr(plan.B), arena.ptr(plan.C)
    g.C_ones = arena.ptr(plan.ones)
    g.bias = arena.ptr(plan.bias)
    g.mask = arena.ptr(plan.mask)
    g.gather_row = SENTINEL if plan.gather_row else None
    g.gather_red = SENTINEL if plan.gather_red else None
    g.M, g.N, g.K = plan.M, plan.N, plan.K
    g.rowA, g.rowC, g.rowM = _aff(plan.rowA), _aff(plan.rowC), _aff(plan.rowM)
    g.redA, g.redB = _aff(plan.redA), _aff(plan.redB)
    g.colB, g.colC, g.colM = _aff(plan.colB), _aff(plan.colC), _aff(plan.colM)
    g.a_dtype = int(plan.a_u8)
    g.ones_row = 1 if plan.ones is not None else 0
    g.ones_stride = plan.ones_stride
    g.bias_mode, g.act, g.mask_mode = plan.bias_mode, plan.act, plan.mask_mode
    g.accumulate = plan.accumulate if plan.splits == 1 else 0      # split-K: the fold accumulates instead
    g.alpha = plan.alpha
    g.batch = plan.batch
    for z in range(4):
        g.batch_offA[z], g.batch_offB[z] = plan.batch_offA[z], plan.batch_offB[z]
        g.batch_offC[z], g.batch_offM[z] = plan.batch_offC[z], plan.batch_offM[z]
    g.splits, g.split_stride = plan.splits, plan.split_stride
    g.vecA, g.vecB, g.tile = plan.vecA, plan.vecB, plan.tile
    g.tc_chunk = plan.tc_chunk
    g.sym = plan.sym
    g.B_image = arena.ptr(plan.b_image)
    return g


def wants_b_image(plan):
    """B operand = network weights (flat parameter buffer) of a GEMM the auto-selection sends to the
    tensor-core engine (mirrors prefer_tc in csrc/gemm.cu)."""
    return (plan.B[0] == "P" and plan.vecA != 2 and plan.M >= 2048 and plan.K >= 64 and plan.tile < 1000
            and not plan.a_u8 == 99)


def forward_splits(M, N, K):
    """Split-K for a forward GEMM whose output grid alone cannot occupy the machine (small batches)."""
    tiles = ((M + 63) // 64) * ((N + 63) // 64)
    ktiles = (K + 15) // 16
    want = (148 + tiles - 1) // tiles
    return int(max(1, min(want, ktiles // 8 if ktiles >= 16 else 1, 32)))


def choose_splits(rows, cols, red):
    """Split the reduction so that a weight-gradient GEMM with a tiny output fills 148 SMs."""
    tiles = ((rows + 127) // 128) * ((cols + 63) // 64 if cols > 32 else 1)
    ktiles = (red + 15) // 16
    want = max(1, (2 * 148 + tiles - 1) // tiles)
    return int(max(1, min(want, ktiles // 8 if ktiles >= 16 else 1, 256)))


class Program(object):
    """An op list plus everything it must keep alive."""

    def __init__(self, device):
        self.device = device
        self.entries = []          # (kind, flags, ctypes struct)
        self.labels = []           # (label or None, flops) per entry -- bench.py's per-op table
        self.flops = 0
        self.plan_names = []
        self._arr = None

    def add(self, kind, desc, flags=0, label=None, flops=0.0):
        self.entries.append((kind, flags, desc))
        self.labels.append((label, flops))
        self._arr = None
        return desc

    def add_plan(self, plan, arena):
        from . import _tcx
        if isinstance(plan, _tcx.TcxOp):
            self.add(plan.kind, plan.make(arena), RUNTIME_IDX if plan.runtime_idx else 0, label=plan.name, flops=plan.flops)
            self.flops += plan.flops
            return
        if isinstance(plan, RawOp):
            d = (plan.struct or GruDesc)()
            for k, ref in plan.fields.items():
                setattr(d, k, arena.ptr(ref))
            for k, v in plan.ints.items():
                setattr(d, k, v)
            if plan.runtime_idx:
                d.idx = SENTINEL
            self.add(plan.kind, d, RUNTIME_IDX if plan.runtime_idx else 0, label=plan.name)
            self.plan_names_all = getattr(self, "plan_names_all", [])
            return
        flags = RUNTIME_IDX if (plan.gather_row or plan.gather_red) else 0
        if plan.b_image is None and wants_b_image(plan):
            # weights feeding a tensor-core GEMM: pre-tile + pre-split them once per step (they change with
            # every optimiser step) so that every CTA fetches its B blocks by TMA instead of gathering them
            bn = 32 if plan.N <= 32 else (64 if plan.N <= 64 else 128)
            floats = plan.batch * ((plan.N + bn - 1) // bn) * ((plan.K + 31) // 32) * 2 * bn * 32
            self._n_images = getattr(self, "_n_images", 0) + 1
            name = "bimg_%x_%d" % (id(self) & 0xFFFFFF, self._n_images)
            arena.ensure(name, floats)
            plan.b_image = (name, 0)
        desc = resolve(plan, arena)
        if plan.b_image is not None:
            self.add(OP_GEMM_PREP_B, desc, 0)
        self.add(OP_GEMM, desc, flags)
        self.plan_names.append(plan.name)
        self.flops += plan.flops
        if plan.splits > 1 and plan.split_dst is not None:
            dst = plan.split_dst
            r = ReduceDesc(arena.ptr((dst[0], dst[1])), arena.ptr(plan.C), dst[2], plan.split_stride, plan.splits,
                           plan.accumulate, arena.ptr(plan.split_bias), plan.N, plan.split_act, plan.split_beta, 0)
            self.add(OP_REDUCE, r)

    def insert(self, index, kind, desc, flags=0, label=None):
        self.entries.insert(index, (kind, flags, desc))
        self.labels.insert(index, (label, 0.0))
        self._arr = None
        return desc

    def index_of(self, label):
        """Index of the first entry carrying `label` (None if absent)."""
        for i, (lab, _) in enumerate(self.labels):
            if lab == label:
                return i
        return None

    def extend(self, other):
        self.entries.extend(other.entries)
        self.labels.extend(other.labels)
        self.flops += other.flops
        self._arr = None

    def finalize(self):
        arr = (Op * len(self.entries))()
        for i, (kind, flags, desc) in enumerate(self.entries):
            arr[i].kind, arr[i].flags = kind, flags
            arr[i].desc = ctypes.addressof(desc)
        self._arr = arr
        return self

    def __len__(self):
        return len(self.entries)

    def n_launches(self):
        return len(self.entries)

    def run(self, idx=None):
        if self._arr is None:
            self.finalize()
        idx_p = _native.dptr(idx, torch.int64, "idx") if idx is not None else c_vp(0)
        with torch.cuda.device(self.device):
            rc = _native.lib().a2cb_run_ops(self._arr, len(self.entries), idx_p, _native.stream_ptr(self.device))
        _native.check(rc, "a2cb_run_ops")


# ---------------------------------------------------------------------------
# parameter layout
# ---------------------------------------------------------------------------
def _pad4(n):
    return (n + 3) // 4 * 4


class ParamLayout(object):
    """Flat fp32 layout of all parameters: each layer is [weight | bias] (so a split-K weight
    gradient and its bias gradient are one contiguous span), segment starts 16 B aligned.
    `entries`: reference state_dict key -> (offset, shape)."""

    def __init__(self, spec):
        self.spec = spec
        self.entries = {}
        self.layers = {}
        off = 0

        def layer(name, wkey, wshape, bkey, bshape):
            nonlocal off
            off = _pad4(off)
            wn = int(math.prod(wshape))
            self.entries[wkey] = (off, tuple(wshape))
            self.entries[bkey] = (off + wn, tuple(bshape))
            self.layers[name] = (off, off + wn)
            off += wn + int(math.prod(bshape))

        H, A = spec.hidden, spec.num_actions
        if spec.cnn:
            C = spec.obs_shape[0]
            layer("conv1", "base.main.0.weight", (32, C, 8, 8), "base.main.0.bias", (32,))
            layer("conv2", "base.main.2.weight", (64, 32, 4, 4), "base.main.2.bias", (64,))
            layer("conv3", "base.main.4.weight", (32, 64, 3, 3), "base.main.4.bias", (32,))
            layer("fc", "base.main.7.weight", (H, 32 * 7 * 7), "base.main.7.bias", (H,))
            gru_in = H
        else:
            gru_in = spec.obs_shape[0]
            d_in = H if spec.recurrent else spec.obs_shape[0]
            layer("actor0", "base.actor.0.weight", (H, d_in), "base.actor.0.bias", (H,))
            layer("actor2", "base.actor.2.weight", (H, H), "base.actor.2.bias", (H,))
            layer("critic0", "base.critic.0.weight", (H, d_in), "base.critic.0.bias", (H,))
            layer("critic2", "base.critic.2.weight", (H, H), "base.critic.2.bias", (H,))
        if spec.recurrent:
            off = _pad4(off)
            self.layers["gru"] = (off, off + 3 * H * gru_in)
            self.entries["base.gru.weight_ih_l0"] = (off, (3 * H, gru_in)); off += 3 * H * gru_in
            self.entries["base.gru.weight_hh_l0"] = (off, (3 * H, H)); off += 3 * H * H
            self.entries["base.gru.bias_ih_l0"] = (off, (3 * H,)); off += 3 * H
            self.entries["base.gru.bias_hh_l0"] = (off, (3 * H,)); off += 3 * H
        # heads: [critic_linear.weight (1,H) | head weight (A,H) | critic_linear.bias | head bias]
        off = _pad4(off)
        hw = "dist.linear" if spec.action_kind == "discrete" else "dist.fc_mean"
        self.entries["base.critic_linear.weight"] = (off, (1, H))
        self.entries[hw + ".weight"] = (off + H, (A, H))
        self.entries["base.critic_linear.bias"] = (off + (1 + A) * H, (1,))
        self.entries[hw + ".bias"] = (off + (1 + A) * H + 1, (A,))
        self.layers["heads"] = (off, off + (1 + A) * H)
        off += (1 + A) * H + 1 + A
        if spec.action_kind == "box":
            off = _pad4(off)
            self.entries["dist.logstd._bias"] = (off, (A, 1))
            self.layers["logstd"] = (off, off)
            off += A
        self.total = _pad4(off)


# ---------------------------------------------------------------------------
# network plans
# ---------------------------------------------------------------------------
class NetBuilder(object):
    """Builds forward / backward plan lists for one policy spec and one batch size."""

    def __init__(self, spec, layout, B, obs_buf, obs_code, gather, tag, seq=None, masks=("masks_in", 0),
                 hxs=("hxs_in", 0), keep_gates=True, allow_tcx=True, full_rows=0, allow_mlp_fused=True, ring=None):
        self.spec, self.L, self.B = spec, layout, B
        # frame ring (RolloutStorage(frame_ring=True)): obs_buf holds uint8 [slots, ring_n, 84, 84] frames stored once,
        # ring = ((arena name, offset) of the per-row valid counts, ring_n)
        self.ring = ring
        # recurrent policies see the batch as a [T, N] sequence, rows time-major (model.py:119-126)
        self.seqT, self.seqN = seq if seq is not None else (1, B)
        assert self.seqT * self.seqN == B
        self.masks_ref, self.hxs_ref, self.keep_gates = masks, hxs, keep_gates
        self.full_rows = full_rows  # rows of the observation source when it is a whole rollout read through a row list
        self.obs = obs_buf          # arena name of the observation source
        self.obs_code = obs_code    # a_dtype of the first layer: 0 raw fp32, 1 uint8/255, 2 fp32/255
        self.gather = gather
        self.t = tag                # arena name prefix for this batch size's buffers
        self.sizes = {}             # arena buffer name -> numel (fp32)
        H, A = spec.hidden, spec.num_actions
        self.zs = 1 + A
        if spec.cnn:
            C = spec.obs_shape[0]
            assert spec.obs_shape[1:] == (84, 84), "CNNBase geometry is the reference's 84x84 (model.py:176-180)"
            self.c1 = P.Conv("conv1", C, 84, 84, 8, 4, 32, in_layout="nchw", needs_dgrad=False)
            self.c2 = P.Conv("conv2", 32, 20, 20, 4, 2, 64)
            self.c3 = P.Conv("conv3", 64, 9, 9, 3, 1, 32)
            self.fc = P.Linear("fc", 1568, H)     # a3 is kept in NCHW-flatten order: no permutation
        self.heads = P.Linear("heads", H, 1 + A)
        # CNN trunk (+ the GRU's input product) on the TMA-fed tensor-core engine (csrc/tcx.cu) unless the exact
        # fp32 SIMT engine was requested; other geometries stay on the affine-indexed engine
        # MLPBase without GRU: both trunks + heads fused into one launch per direction (csrc/mlp_fused.cu)
        # (KFAC needs every layer's input and output gradient in memory: it asks for the layer-by-layer plans)
        self.mlp_fused = (allow_mlp_fused and not spec.cnn and not spec.recurrent and os.environ.get("A2CB_MLP_FUSED", "1") != "0"
                          and bool(_native.lib().a2cb_mlp_fused_supported(int(spec.obs_shape[0]), int(H), 1 + A)))
        self.tcx = None
        if (allow_tcx and spec.cnn and spec.obs_shape[0] == 4 and H % 256 == 0 and obs_code in (1, 2)
                and _native.lib().a2cb_get_gemm_mode() != 0 and os.environ.get("A2CB_TCX", "1") != "0"):
            from . import _tcx
            self.tcx = _tcx.CnnNet(self)
        if ring is not None and self.tcx is None:
            raise _native.NativeError("a frame-ring rollout is read by the TMA-fed engine only (4 x 84 x 84 uint8 stacks, "
                                      "hidden size a multiple of 256, gemm mode != 0)")

    def n(self, name):
        return self.t + name

    def buf(self, name, numel):
        self.sizes[self.n(name)] = int(numel)
        return (self.n(name), 0)

    def wb(self, layer):
        w, b = self.L.layers[layer]
        return ("P", w), ("P", b)

    def gwb(self, layer):
        w, b = self.L.layers[layer]
        return ("G", w), ("G", b)

    # ---- CNN ----
    def cnn_forward(self):
        B, s = self.B, self.spec
        c1, c2, c3 = self.c1, self.c2, self.c3
        a1, a2, a3 = self.buf("a1", B * c1.out_row), self.buf("a2", B * c2.out_row), self.buf("a3", B * c3.out_row)
        if self.tcx is not None:
            self.buf("z", B * self.zs)
            return self.tcx.forward()
        h, z = self.buf("h", B * s.hidden), self.buf("z", B * self.zs)
        sp = forward_splits(B, s.hidden, 1568)
        part = self.buf("part_fc_fwd", sp * B * s.hidden) if sp > 1 else None
        plans = [
            P.conv_forward(c1, B, (self.obs, 0), *self.wb("conv1"), a1, a_u8=self.obs_code, gather=self.gather),
            P.conv_forward(c2, B, a1, *self.wb("conv2"), a2),
            P.conv_forward(c3, B, a2, *self.wb("conv3"), a3, out_layout="nchw"),
            P.linear_forward(self.fc, B, a3, *self.wb("fc"), h, act=P.ACT_RELU, splits=sp, partial=part),
        ]
        return plans, h

    def cnn_backward(self, feat_grad, feat_grad_lo=None):
        """feat_grad: gradient w.r.t. the trunk output h, already masked by relu' (feat_grad_lo: its lo plane when
        the TMA-fed engine produced it)."""
        B = self.B
        if self.tcx is not None:
            from . import _tcx
            if feat_grad_lo is not None:
                return self.tcx.backward(feat_grad, feat_grad_lo)
            lo = self.buf("feat_grad_lo", B * self.spec.hidden)
            return [_tcx.lo_op("lo_plane", feat_grad, lo, B * self.spec.hidden)] + self.tcx.backward(feat_grad, lo)
        c1, c2, c3 = self.c1, self.c2, self.c3
        a1, a2, a3 = (self.n("a1"), 0), (self.n("a2"), 0), (self.n("a3"), 0)
        g3, g2, g1 = self.buf("g3", B * c3.gout_row), self.buf("g2", B * c2.gout_row), self.buf("g1", B * c1.gout_row)
        plans = [
            P.linear_wgrad(self.fc, B, a3, feat_grad, *self.gwb("fc")),
            P.linear_dgrad(self.fc, B, feat_grad, ("P", self.L.layers["fc"][0]), g3, mask=a3, mask_mode=1,
                           row_out=P.plain(c3.gout_row), col_out=P.aff(49, 7, 1, c3.pw * 32, 32),
                           gx_base=c3.gout_base, mask_row=P.plain(1568), mask_col=P.plain(1)),
        ]
        for cv, src, g, gprev, prev, a_prev in ((c3, a2, g3, g2, c2, a2), (c2, a1, g2, g1, c1, a1)):
            plans += self.conv_wgrad_ops(cv, src, g)
            plans.append(P.conv_dgrad(cv, B, g, ("P", self.L.layers[cv.name][0]), gprev,
                                      (prev.gout_row, prev.pw, prev.gout_base), mask=a_prev))
        plans += self.conv_wgrad_ops(c1, (self.obs, 0), g1, a_u8=self.obs_code, gather=self.gather)
        return plans

    def conv_wgrad_ops(self, cv, src, g, a_u8=False, gather=False):
        """Weight (+ bias) gradient of one conv layer.  The bias gradient is normally the logical ones row
        of the same GEMM; when the patch size K is a multiple of the 128-row tile that row would cost a
        whole extra tile row of CTAs (conv1: +50%), so it is computed as a pixel sum + batch fold instead."""
        B = self.B
        wref, bref = self.gwb(cv.name)
        fused_bias = cv.K % 128 != 0
        rows = cv.K + (1 if fused_bias else 0)
        sp = choose_splits(rows, cv.cout, B * cv.oh * cv.ow)
        span = cv.cout * cv.K + (cv.cout if fused_bias else 0)
        part = self.buf("part_" + cv.name, sp * span) if sp > 1 else None
        ops = [P.conv_wgrad(cv, B, src, g, wref, bref if fused_bias else None, a_u8=a_u8, gather=gather,
                            splits=sp, partial=part)]
        if not fused_bias:
            s = self.buf("bsum_" + cv.name, B * cv.cout)
            ops.append(RawOp(OP_SPATIAL_SUM, {"out": s, "in_": (g[0], g[1] + cv.gout_base)},
                             dict(img_stride=cv.gout_row, pitch=cv.pw * cv.cout, B=B, P=cv.oh * cv.ow, C=cv.cout,
                                  ow=cv.ow), False, cv.name + ".bias_psum", SpatialSumDesc))
            ops.append(RawOp(OP_REDUCE, {"dst": bref, "src": s},
                             dict(count=cv.cout, stride=cv.cout, splits=B, accumulate=0, ncols=0, act=0, beta=1.0),
                             False, cv.name + ".bias_fold", ReduceDesc))
        return ops

    # ---- MLP ----
    def mlp_forward(self, x=None, d_in=None):
        B, s = self.B, self.spec
        H = s.hidden
        from_obs = x is None
        d = d_in if d_in is not None else s.obs_shape[0]
        x = x if x is not None else (self.obs, 0)
        ha1, ha2 = self.buf("ha1", B * H), self.buf("ha2", B * H)
        hc1, hc2 = self.buf("hc1", B * H), self.buf("hc2", B * H)
        z = self.buf("z", B * self.zs)
        l0a, l2a = P.Linear("actor0", d, H), P.Linear("actor2", H, H)
        l0c, l2c = P.Linear("critic0", d, H), P.Linear("critic2", H, H)
        self.mlp_layers = (l0a, l2a, l0c, l2c)
        p0a = P.linear_forward(l0a, B, x, *self.wb("actor0"), ha1, act=P.ACT_TANH)
        p0c = P.linear_forward(l0c, B, x, *self.wb("critic0"), hc1, act=P.ACT_TANH)
        for p in (p0a, p0c):
            p.gather_row = self.gather and from_obs
            p.vecA = 1 if d % 4 == 0 else 0
        plans = [p0c, P.linear_forward(l2c, B, hc1, *self.wb("critic2"), hc2, act=P.ACT_TANH),
                 p0a, P.linear_forward(l2a, B, ha1, *self.wb("actor2"), ha2, act=P.ACT_TANH)]
        return plans, (hc2, ha2)

    def mlp_fused_fields(self):
        B, H = self.B, self.spec.hidden
        lay, L = self.L.layers, self.L
        hw, hb = lay["heads"]
        span = hb + self.zs
        ints = dict(off_a0w=lay["actor0"][0], off_a0b=lay["actor0"][1], off_a2w=lay["actor2"][0], off_a2b=lay["actor2"][1],
                    off_c0w=lay["critic0"][0], off_c0b=lay["critic0"][1], off_c2w=lay["critic2"][0], off_c2b=lay["critic2"][1],
                    off_hw=hw, off_hb=hb, span=span, B=B, d_in=self.spec.obs_shape[0], H=H, zs=self.zs)
        fields = dict(P=("P", 0), G=("G", 0), obs=(self.obs, 0), acts=self.buf("mlp_acts", 4 * B * H),
                      z=self.buf("z", B * self.zs))
        return fields, ints, span

    def mlp_fused_op(self, backward):
        fields, ints, span = self.mlp_fused_fields()
        if backward:
            fields["dz"] = self.buf("dz", self.B * self.zs)
            ncta = (self.B + 63) // 64
            if ncta > 1:
                fields["partial"] = self.buf("mlp_gpart", ncta * span)
        return RawOp(OP_MLP_BWD if backward else OP_MLP_FWD, fields, ints, self.gather,
                     "mlp.fused_bwd" if backward else "mlp.fused_fwd", MlpDesc)

    def mlp_fused_backward(self):
        op = self.mlp_fused_op(True)
        ncta = (self.B + 63) // 64
        if ncta == 1:
            return [op]
        span = op.ints["span"]
        return [op, RawOp(OP_REDUCE, {"dst": ("G", 0), "src": op.fields["partial"]},
                          dict(count=span, stride=span, splits=ncta, accumulate=0, ncols=0, act=0, beta=1.0), False,
                          "mlp.fold", ReduceDesc)]

    def mlp_heads_forward(self, hc2, ha2):
        B, H, A = self.B, self.spec.hidden, self.spec.num_actions
        w0, b0 = self.L.layers["heads"]
        z = (self.n("z"), 0)
        crit = P.Linear("critic_linear", H, 1)
        mean = P.Linear("head", H, A)
        return [P.linear_forward(crit, B, hc2, ("P", w0), ("P", b0), z, y_stride=self.zs),
                P.linear_forward(mean, B, ha2, ("P", w0 + H), ("P", b0 + 1), (z[0], 1), y_stride=self.zs)]

    def mlp_backward(self, x=None, d_in=None, gx=None):
        """gx: optional buffer receiving the gradient w.r.t. the trunk input (recurrent MLP)."""
        B, H, A = self.B, self.spec.hidden, self.spec.num_actions
        from_obs = x is None
        d = d_in if d_in is not None else self.spec.obs_shape[0]
        l0a, l2a, l0c, l2c = self.mlp_layers
        w0, b0 = self.L.layers["heads"]
        dz = (self.n("dz"), 0)
        ha1, ha2, hc1, hc2 = [(self.n(k), 0) for k in ("ha1", "ha2", "hc1", "hc2")]
        gc2, gc1 = self.buf("gc2", B * H), self.buf("gc1", B * H)
        ga2, ga1 = self.buf("ga2", B * H), self.buf("ga1", B * H)
        crit = P.Linear("critic_linear", H, 1)
        mean = P.Linear("head", H, A)
        x = x if x is not None else (self.obs, 0)
        plans = [
            P.linear_wgrad(crit, B, hc2, dz, ("G", w0), ("G", b0), gy_stride=self.zs),
            P.linear_dgrad(crit, B, dz, ("P", w0), gc2, mask=hc2, mask_mode=2, gy_stride=self.zs),
            P.linear_wgrad(l2c, B, hc1, gc2, *self.gwb("critic2")),
            P.linear_dgrad(l2c, B, gc2, ("P", self.L.layers["critic2"][0]), gc1, mask=hc1, mask_mode=2),
            P.linear_wgrad(l0c, B, x, gc1, *self.gwb("critic0")),
            P.linear_wgrad(mean, B, ha2, (dz[0], 1), ("G", w0 + H), ("G", b0 + 1), gy_stride=self.zs),
            P.linear_dgrad(mean, B, (dz[0], 1), ("P", w0 + H), ga2, mask=ha2, mask_mode=2, gy_stride=self.zs),
            P.linear_wgrad(l2a, B, ha1, ga2, *self.gwb("actor2")),
            P.linear_dgrad(l2a, B, ga2, ("P", self.L.layers["actor2"][0]), ga1, mask=ha1, mask_mode=2),
            P.linear_wgrad(l0a, B, x, ga1, *self.gwb("actor0")),
        ]
        for p in (plans[4], plans[9]):
            p.gather_red = self.gather and from_obs
            if self.gather and from_obs:
                p.vecA = 0      # gathered rows: row-contiguity of x across the batch index does not apply
        if gx is not None:
            plans.append(P.linear_dgrad(l0c, B, gc1, ("P", self.L.layers["critic0"][0]), gx))
            plans.append(P.linear_dgrad(l0a, B, ga1, ("P", self.L.layers["actor0"][0]), gx, accumulate=1))
        return plans

    # ---- GRU (model.py:111-166) ----
    def gru_forward(self, x, d_in):
        """x: (buffer, offset) of the [B, d_in] input rows.  Returns (ops, output ref)."""
        B, H, T, N = self.B, self.spec.hidden, self.seqT, self.seqN
        e = self.L.entries
        wih, whh = ("P", e["base.gru.weight_ih_l0"][0]), ("P", e["base.gru.weight_hh_l0"][0])
        bih, bhh = ("P", e["base.gru.bias_ih_l0"][0]), ("P", e["base.gru.bias_hh_l0"][0])
        gi, gh = self.buf("gru_gi", B * 3 * H), self.buf("gru_gh", B * 3 * H)
        hm, out = self.buf("gru_hm", B * H), self.buf("gru_out", B * H)
        fields = dict(gi=gi, gh=gh, hm=hm, masks=self.masks_ref, out=out)
        if self.keep_gates:
            fields.update(r=self.buf("gru_r", B * H), z=self.buf("gru_z", B * H), n=self.buf("gru_n", B * H))
        self.gru_in = P.Linear("gru_ih", d_in, 3 * H)
        self.gru_hh = P.Linear("gru_hh", H, 3 * H)
        self.gru_persistent = gru_seq_supported(N, H)
        in_prod = [P.linear_forward(self.gru_in, B, x, wih, bih, gi)]
        if self.tcx is not None and d_in == H:
            in_prod = self.tcx.gru_ih_forward(gi)
        if self.gru_persistent:
            # whole recurrence in one cooperative launch (csrc/gru_persistent.cu)
            f2 = dict(fields, h0=self.hxs_ref, W_hh=whh, b_hh=bhh)
            return in_prod + [
                    RawOp(OP_GRU_SEQ_FWD, f2, dict(T=T, N=N, H=H, t=0), self.gather, "gru.seq_fwd", GruSeqDesc)], out
        ops = in_prod + [
               RawOp(OP_GRU_INIT, dict(hm=hm, h0=self.hxs_ref, masks=self.masks_ref), dict(T=T, N=N, H=H, t=0),
                     self.gather, "gru.init")]
        sp = forward_splits(N, 3 * H, H)          # the per-step GEMM is tiny (M = environments): split K
        part = self.buf("part_gru_hh_fwd", sp * N * 3 * H) if sp > 1 else None
        ints = dict(T=T, N=N, H=H)
        if sp > 1:      # the gate kernel folds the split-K partials itself (no separate fold launch per step)
            fields.update(gh_part=part, gh_bias=bhh)
            ints.update(gh_splits=sp)
        for t in range(T):
            p = P.linear_forward(self.gru_hh, N, (hm[0], t * N * H), whh, bhh, (gh[0], t * N * 3 * H),
                                 splits=sp, partial=part)
            p.name = "gru_hh.fwd"
            if sp > 1:
                p.split_dst = None
            ops.append(p)
            ops.append(RawOp(OP_GRU_FWD, fields, dict(ints, t=t), self.gather, "gru.gates"))
        return ops, out

    def gru_backward(self, x, d_in, dout, dx=None, dx_mask=None):
        """dout: gradient w.r.t. the GRU outputs [B, H].  Writes the GRU parameter gradients and, when
        dx is given, the gradient w.r.t. the input rows (masked by relu' of dx_mask)."""
        B, H, T, N = self.B, self.spec.hidden, self.seqT, self.seqN
        e = self.L.entries
        whh, wih = ("P", e["base.gru.weight_hh_l0"][0]), ("P", e["base.gru.weight_ih_l0"][0])
        dgi, dgh = self.buf("gru_dgi", B * 3 * H), self.buf("gru_dgh", B * 3 * H)
        dcarry = self.buf("gru_dcarry", N * H)
        names = ("gi", "gh", "hm", "out", "r", "z", "n")
        fields = {k: (self.n("gru_" + k), 0) for k in names}
        fields.update(masks=self.masks_ref, dout=dout, dgi=dgi, dgh=dgh, dcarry=dcarry)
        ops = []
        if getattr(self, "gru_persistent", False):
            f2 = dict(fields, W_hh=whh)          # dcarry doubles as the kernel's carry workspace for N > 64
            ops.append(RawOp(OP_GRU_SEQ_BWD, f2, dict(T=T, N=N, H=H, t=0), self.gather, "gru.seq_bwd", GruSeqDesc))
        sp = forward_splits(N, H, 3 * H)
        part = self.buf("part_gru_hh_dgrad", sp * N * H) if sp > 1 else None
        ints = dict(T=T, N=N, H=H)
        if sp > 1:
            fields.update(dc_part=part)
            ints.update(dc_splits=sp)
        for t in range(T - 1, -1, -1):
            if getattr(self, "gru_persistent", False):
                break
            ops.append(RawOp(OP_GRU_BWD, fields, dict(ints, t=t), self.gather, "gru.gates_bwd"))
            if t > 0:
                p = P.linear_dgrad(self.gru_hh, N, (dgh[0], t * N * 3 * H), whh, dcarry, accumulate=1,
                                   splits=sp, partial=part)
                if sp > 1:
                    p.split_dst = None      # folded by the next (earlier-in-time) backward gate kernel
                p.name = "gru_hh.dgrad"
                ops.append(p)
        hm = (self.n("gru_hm"), 0)
        if self.tcx is not None and d_in == H and dx is not None and dx_mask is not None:
            return ops + self.tcx.gru_backward(dgi, dgh, hm, dx, (dx[0] + "_lo", dx[1]))
        ops.append(P.linear_wgrad(self.gru_hh, B, hm, dgh, ("G", e["base.gru.weight_hh_l0"][0]),
                                  ("G", e["base.gru.bias_hh_l0"][0])))
        ops.append(P.linear_wgrad(self.gru_in, B, x, dgi, ("G", e["base.gru.weight_ih_l0"][0]),
                                  ("G", e["base.gru.bias_ih_l0"][0])))
        if dx is not None:
            ops.append(P.linear_dgrad(self.gru_in, B, dgi, wih, dx, mask=dx_mask, mask_mode=1 if dx_mask else 0))
        return ops

    # ---- heads (CNN: one [1+A, H] matrix on the shared feature) ----
    def cnn_heads_forward(self, h):
        w, b = self.wb("heads")
        if self.tcx is not None and self.zs <= 8:
            # [1 + A, H] heads: one streaming pass over the features (a warp per row)
            return [_tcx.narrow_fwd_op("heads.fwd", (self.n("z"), 0), h, w, b, self.B, self.spec.hidden, self.zs, self.zs)]
        sp = forward_splits(self.B, self.zs, self.spec.hidden)
        part = self.buf("part_heads_fwd", sp * self.B * self.zs) if sp > 1 else None
        return [P.linear_forward(self.heads, self.B, h, w, b, (self.n("z"), 0), splits=sp, partial=part)]

    def cnn_heads_backward(self, h, relu_mask=True):
        B, H = self.B, self.spec.hidden
        dz = self.buf("dz", B * self.zs)
        gh = self.buf("gh", B * H)
        gw, gb = self.gwb("heads")
        if self.tcx is not None and self.zs <= 8:
            # [1 + A, H] weight gradient + bias gradient (contiguous in the flat layout): one streaming pass
            wg = _tcx.narrow_wgrad_ops(self, "heads", h, dz, B, H, self.zs, self.zs, gw)
        else:
            wg = [P.linear_wgrad(self.heads, B, h, dz, gw, gb)]
        if self.tcx is not None and self.zs <= 8:
            # gradient w.r.t. the features, relu' mask and lo plane in the same pass
            self.gh_lo = self.buf("gh_lo", B * H) if relu_mask else None      # the GRU consumes the plain plane only
            return wg + [_tcx.narrow_dgrad_op("heads.dgrad", gh, self.gh_lo, dz, ("P", self.L.layers["heads"][0]),
                                              h if relu_mask else None, B, H, self.zs, self.zs)], gh
        return wg + [
                P.linear_dgrad(self.heads, B, dz, ("P", self.L.layers["heads"][0]), gh,
                               mask=h if relu_mask else None, mask_mode=1 if relu_mask else 0)], gh

    def forward_plans(self):
        s = self.spec
        if s.cnn:
            plans, h = self.cnn_forward()
            if s.recurrent:
                gops, feat = self.gru_forward(h, s.hidden)
                return plans + gops + self.cnn_heads_forward(feat)
            return plans + self.cnn_heads_forward(h)
        if s.recurrent:
            gops, feat = self.gru_forward((self.obs, 0), s.obs_shape[0])
            # the GRU reads the (possibly gathered) observation rows itself; the trunks read its output
            gops[0].gather_row = self.gather
            gops[0].vecA = 1 if s.obs_shape[0] % 4 == 0 else 0
            plans, (hc2, ha2) = self.mlp_forward(x=feat, d_in=s.hidden)
            return gops + plans + self.mlp_heads_forward(hc2, ha2)
        if self.mlp_fused:
            return [self.mlp_fused_op(False)]
        plans, (hc2, ha2) = self.mlp_forward()
        return plans + self.mlp_heads_forward(hc2, ha2)

    def backward_plans(self):
        s = self.spec
        if s.cnn:
            h = (self.n("h"), 0)
            if s.recurrent:
                feat = (self.n("gru_out"), 0)
                hp, gfeat = self.cnn_heads_backward(feat, relu_mask=False)
                gx = self.buf("gru_dx", self.B * s.hidden)
                gx_lo = self.buf("gru_dx_lo", self.B * s.hidden) if self.tcx is not None else None
                return hp + self.gru_backward(h, s.hidden, gfeat, dx=gx, dx_mask=h) + self.cnn_backward(gx, gx_lo)
            hp, gh = self.cnn_heads_backward(h)
            return hp + self.cnn_backward(gh, getattr(self, "gh_lo", None))
        self.buf("dz", self.B * self.zs)
        if s.recurrent:
            feat = (self.n("gru_out"), 0)
            gfeat = self.buf("gru_dout", self.B * s.hidden)
            plans = self.mlp_backward(x=feat, d_in=s.hidden, gx=gfeat)
            gb = self.gru_backward((self.obs, 0), s.obs_shape[0], gfeat)
            w = gb[-1]                       # dW_ih reads the (possibly gathered) observation rows
            w.gather_red = self.gather
            w.vecA = 0 if self.gather else w.vecA
            return plans + gb
        if self.mlp_fused:
            return self.mlp_fused_backward()
        return self.mlp_backward()


# ---------------------------------------------------------------------------
# runtime: caches programs per batch shape and runs them
# ---------------------------------------------------------------------------
def derive_sampling_seed(cpu_seed):
    """Philox key of the sampling stream from torch's CPU seed, WITHOUT consuming the generator (the minibatch
    permutations of update() must stay the reference's)."""
    return (int(cpu_seed) * 0x9E3779B97F4A7C15 + 0x632BE59BD9B4E019) & 0x7FFFFFFFFFFFFFFF


class PolicyRuntime(object):
    MAX_CACHE = 96
    STAGE_BYTES = 64 << 20

    def __init__(self, spec, layout, flat, flat_grad):
        self.spec, self.L = spec, layout
        self.dev = flat.device
        if self.dev.type != "cuda":
            raise _native.NativeError("the policy lives on %s: move it to a CUDA device "
                                      "(the update path has no CPU fallback)" % self.dev)
        _native.lib()
        self.arena = Arena(self.dev)
        self.arena.bind("P", flat)
        self.arena.bind("G", flat_grad)
        self.cache = {}
        self.forward_token = 0
        self.keep = {}
        self.zs = 1 + spec.num_actions
        self.rng_provider = None     # callable -> device int64[4] {seed, offset, ticket, row_base}

    # -- helpers ------------------------------------------------------------
    def obs_code(self, obs):
        if obs.dtype == torch.uint8:
            if not self.spec.cnn:
                raise _native.NativeError("uint8 observations are only meaningful for the CNN trunk")
            return 1
        if obs.dtype != torch.float32:
            raise _native.NativeError("observations must be float32 or uint8 (got %s)" % obs.dtype)
        return 2 if self.spec.cnn else 0

    def _remember(self, key, value):
        if len(self.cache) >= self.MAX_CACHE:
            self.cache.pop(next(iter(self.cache)))
        self.cache[key] = value
        return value

    def _alloc(self, builder):
        for name, numel in builder.sizes.items():
            self.arena.ensure(name, numel)

    def loss_scratch(self, B):
        n = int(_native.lib().a2cb_loss_scratch_floats(B))
        return self.arena.ensure("loss_scratch_%d" % B, n)

    def logstd_ref(self):
        if self.spec.action_kind != "box":
            return None
        return ("P", self.L.entries["dist.logstd._bias"][0])

    # -- actor step ---------------------------------------------------------
    def rng_state(self):
        """Device {seed, offset, ticket, row_base} of the sampling stream (a2cb_sample_t.rng).  Owned by the Policy
        (`Policy.sampling_state`), so that it survives re-flattening, device moves and pickling; a runtime built
        without a policy (tests) keeps its own."""
        if self.rng_provider is not None:
            return self.rng_provider()
        st = getattr(self, "_rng_state", None)
        if st is None:
            st = self._rng_state = torch.tensor([derive_sampling_seed(torch.initial_seed()), 0, 0, 0],
                                                dtype=torch.int64, device=self.dev)
        return st

    def sample(self, z, action, value, logp, deterministic):
        """Sampling head alone on head outputs z [B, 1 + A] (contiguous)."""
        d = SampleDesc()
        d.z, d.logstd = z.data_ptr(), self.arena.ptr(self.logstd_ref())
        d.value_out, d.action_out, d.logp_out = value.data_ptr(), action.data_ptr(), logp.data_ptr()
        d.rng = self.rng_state().data_ptr()
        d.B, d.A, d.z_stride, d.dist = z.shape[0], self.spec.num_actions, z.stride(0), 0 if self.spec.action_kind == "discrete" else 1
        d.deterministic = int(bool(deterministic))
        with torch.cuda.device(self.dev):
            rc = _native.lib().a2cb_sample_actions(ctypes.byref(d), _native.stream_ptr(self.dev))
        _native.check(rc, "a2cb_sample_actions")

    def act_plan(self, B, obs_dtype):
        key = ("act", B, obs_dtype)
        plan = self.cache.get(key)
        if plan is not None:
            return plan
        s, A = self.spec, self.spec.num_actions
        tag = "a%d_" % B
        n_obs = B * int(math.prod(s.obs_shape))
        t = {"obs": torch.zeros(n_obs, dtype=obs_dtype, device=self.dev),
             "value": torch.zeros(B, device=self.dev), "logp": torch.zeros(B, device=self.dev),
             "action": (torch.zeros(B, dtype=torch.int64, device=self.dev) if s.action_kind == "discrete"
                        else torch.zeros(B * A, device=self.dev))}
        code = self.obs_code(t["obs"])
        self.arena.bind("obs_in", t["obs"])
        seq = None
        if s.recurrent:
            t["hxs"] = torch.zeros(B * s.hidden, device=self.dev)
            t["masks"] = torch.ones(B, device=self.dev)
            self.arena.bind("hxs_in", t["hxs"])
            self.arena.bind("masks_in", t["masks"])
            seq = (1, B)
        nb = NetBuilder(s, self.L, B, "obs_in", code, False, tag, seq=seq, keep_gates=not s.recurrent)
        plans = nb.forward_plans()
        if nb.tcx is not None:
            plans, = _tcx.hoist_preps([plans], tag + "f")
        self._alloc(nb)
        if s.recurrent:
            t["hout"] = self.arena.bufs[tag + "gru_out"]
            assert t["hout"].numel() == B * s.hidden
        prog = Program(self.dev)
        for p in plans:
            prog.add_plan(p, self.arena)
        d = SampleDesc()
        d.z, d.logstd = self.arena.ptr((tag + "z", 0)), self.arena.ptr(self.logstd_ref())
        d.value_out, d.action_out, d.logp_out = t["value"].data_ptr(), t["action"].data_ptr(), t["logp"].data_ptr()
        d.rng = self.rng_state().data_ptr()
        d.B, d.A, d.z_stride, d.dist = B, A, self.zs, 0 if s.action_kind == "discrete" else 1
        prog.add(OP_SAMPLE, d)
        prog.finalize()
        sites = {}
        for role, ten in t.items():
            lo = ten.data_ptr()
            found = []
            for _, _, desc in prog.entries:
                _pointer_sites(desc, lo, lo + ten.numel() * ten.element_size(), found)
            assert found, "actor program: nothing reads / writes the %s placeholder" % role
            sites[role] = found
        return self._remember(key, ActPlan(prog, d, t, sites))

    # -- forward only -------------------------------------------------------
    def _bind_inputs(self, obs, hxs, masks):
        """Binds the caller's tensors into the arena; returns (B, seq, key part)."""
        if not obs.is_cuda:
            raise _native.NativeError("observations must be CUDA tensors; there is no CPU fallback")
        obs = obs.contiguous()
        B = obs.shape[0]
        if tuple(obs.shape[1:]) != self.spec.obs_shape:
            raise ValueError("expected observations of shape [B, %s], got %s" % (self.spec.obs_shape, tuple(obs.shape)))
        # Actor-side calls pass a different slice of the rollout buffer every step (`rollouts.obs[step]`,
        # reference main.py:117): inputs up to STAGE_BYTES are copied into arena-owned buffers so that ONE cached
        # program per batch size serves every step; larger batches are read in place (program keyed by address).
        if obs.numel() * obs.element_size() <= self.STAGE_BYTES:
            obs = self.arena.stage("obs_in_%d_%s" % (B, obs.dtype), obs).view(obs.shape)
        self.arena.bind("obs_in", obs.view(-1))
        self.keep["obs"] = obs
        seq, extra = None, ()
        if self.spec.recurrent:
            N = hxs.shape[0]
            if B % N != 0:
                raise ValueError("batch of %d rows is not a whole number of %d-environment steps" % (B, N))
            hx = hxs.detach().to(torch.float32).contiguous()
            mk = masks.detach().to(torch.float32).contiguous().view(-1)
            if not (hx.is_cuda and mk.is_cuda):
                raise _native.NativeError("recurrent state and masks must be CUDA tensors")
            if hx.numel() * 4 <= self.STAGE_BYTES:
                hx = self.arena.stage("hxs_in_%d" % hx.numel(), hx).view(hx.shape)
                mk = self.arena.stage("masks_in_%d" % mk.numel(), mk)
            self.arena.bind("hxs_in", hx.view(-1))
            self.arena.bind("masks_in", mk)
            self.keep["hxs"], self.keep["masks"] = hx, mk
            seq, extra = (B // N, N), (hx.data_ptr(), mk.data_ptr(), N)
        return obs, B, seq, extra

    def forward(self, obs, hxs, masks, keep_for_backward=False):
        obs, B, seq, extra = self._bind_inputs(obs, hxs, masks)
        code = self.obs_code(obs)
        key = ("fwd", B, obs.data_ptr(), code, bool(keep_for_backward)) + extra
        ent = self.cache.get(key)
        if ent is None:
            nb = NetBuilder(self.spec, self.L, B, "obs_in", code, False, "b%d_" % B, seq=seq,
                            keep_gates=keep_for_backward or not self.spec.recurrent)
            plans = nb.forward_plans()
            if nb.tcx is not None:
                plans, = _tcx.hoist_preps([plans], "b%d_f" % B)
            self._alloc(nb)
            prog = Program(self.dev)
            for p in plans:
                prog.add_plan(p, self.arena)
            ent = self._remember(key, (prog.finalize(), nb))
        prog, nb = ent
        prog.run()
        self.forward_token += 1
        z = self.arena.bufs["b%d_z" % B][:B * self.zs].view(B, self.zs)
        out_h = hxs
        if self.spec.recurrent:
            T, N = seq
            H = self.spec.hidden
            out_h = self.arena.bufs["b%d_gru_out" % B][(T - 1) * N * H:T * N * H].view(N, H).clone()
        return {"z": z, "hxs": out_h, "builder": nb, "seq": seq}

    def _loss_desc(self, B, tag, mode, actions, **kw):
        s = self.spec
        d = LossDesc()
        d.z = self.arena.ptr((tag + "z", 0))
        d.logstd = self.arena.ptr(self.logstd_ref())
        d.actions = actions.data_ptr()
        d.act_stride = actions.shape[-1]
        d.B, d.A = B, s.num_actions
        d.z_stride = d.dz_stride = self.zs
        d.dlogstd_stride = 1
        d.dist = 0 if s.action_kind == "discrete" else 1
        d.mode = mode
        d.inv_B = 1.0 / B
        d.n_valid = -1
        d.scratch = self.loss_scratch(B).data_ptr()
        for k, v in kw.items():
            setattr(d, k, v)
        return d

    def _check_actions(self, B, action):
        s = self.spec
        want = torch.int64 if s.action_kind == "discrete" else torch.float32
        if not action.is_cuda:
            raise _native.NativeError("actions must be CUDA tensors")
        a = action.to(want).contiguous().view(B, -1)
        if s.action_kind == "box" and a.shape[1] != s.num_actions:
            raise ValueError("action dimension mismatch")
        return a

    def head_stats(self, B, action):
        """log-prob [B] and mean entropy (0-dim) of `action` under the last forward of batch B."""
        a = self._check_actions(B, action)
        out = self.arena.ensure("b%d_stats" % B, 2 * B + 4)
        d = self._loss_desc(B, "b%d_" % B, 3, a, logp_out=out.data_ptr(), loss_out=out.data_ptr() + 8 * B)
        self.keep["stats_actions"] = a
        with torch.cuda.device(self.dev):
            rc = _native.lib().a2cb_loss_epilogue(ctypes.byref(d), _native.stream_ptr(self.dev))
        _native.check(rc, "a2cb_loss_epilogue")
        return {"logp": out[:B], "entropy": out[2 * B + 2]}

    # -- VJP for the autograd bridge -----------------------------------------
    def backward_vjp(self, B, action, g_value, g_logp, g_ent):
        a = self._check_actions(B, action)
        tag = "b%d_" % B
        obs = self.keep["obs"]
        code = self.obs_code(obs)
        seq, extra = None, ()
        if self.spec.recurrent:
            hx, mk = self.keep["hxs"], self.keep["masks"]
            seq, extra = (B // hx.shape[0], hx.shape[0]), (hx.data_ptr(), mk.data_ptr(), hx.shape[0])
        key = ("bwd", B, obs.data_ptr(), code) + extra
        ent = self.cache.get(key)
        self.arena.bind("obs_in", obs.view(-1))
        if ent is None:
            nb = NetBuilder(self.spec, self.L, B, "obs_in", code, False, tag, seq=seq)
            nb.forward_plans()
            plans = nb.backward_plans()
            if nb.tcx is not None:
                plans, = _tcx.hoist_preps([plans], tag + "b")
            self._alloc(nb)
            prog = Program(self.dev)
            for p in plans:
                prog.add_plan(p, self.arena)
            ent = self._remember(key, prog.finalize())
        prog = ent
        d = self._loss_desc(B, tag, 4, a, old_value=g_value.data_ptr(), old_logp=g_logp.data_ptr(),
                            noise=g_ent.data_ptr(), dz=self.arena.ptr((tag + "dz", 0)))
        if self.spec.action_kind == "box":
            d.dlogstd = self.arena.ptr(("G", self.L.entries["dist.logstd._bias"][0]))
        with torch.cuda.device(self.dev):
            rc = _native.lib().a2cb_loss_epilogue(ctypes.byref(d), _native.stream_ptr(self.dev))
        _native.check(rc, "a2cb_loss_epilogue")
        prog.run()
        return self.arena.bufs["G"]

    # -- fused training step --------------------------------------------------
    def train_program(self, storage, mbs, mode, hyper, adv=None, noise=None, loss_accum=None, use_idx=True,
                      with_backward=True, seq=None, tag=None):
        """Op list of one minibatch step reading rows of `storage` through the runtime index list:
        forward -> loss epilogue (mode) -> backward.  Optimiser ops are appended by the caller.
        `tag`: arena name prefix of the program's buffers (default: one set per minibatch size).  Programs of
        different sizes that never run concurrently may share a tag -- build the LARGEST first, since growing a
        buffer re-allocates it under the programs that already point at it."""
        s = self.spec
        T, N = storage.rewards.shape[0], storage.rewards.shape[1]
        ring = None
        if getattr(storage, "frame_ring", False) and not use_idx:
            # full-batch programs (A2C): the stacks are materialised into an arena buffer by the caller before every
            # run (`stage_ring_batch`); the batches of these algorithms are a few hundred rows
            code = 1
            self.stage_ring_batch(storage)
        elif getattr(storage, "frame_ring", False):
            # every frame stored once: the frame re-layout kernel assembles the 4-stacks from the ring
            if tuple(storage.obs.shape[2:]) != s.obs_shape:
                raise _native.NativeError("frame ring: observation shape not supported")
            code = 1
            self.arena.bind("obs_st", storage.frames.view(-1))
            self.arena.bind("obs_valid", storage.stack_valid.view(-1))
            ring = (("obs_valid", 0), N)
        else:
            obs_flat = storage.obs[:-1].view(T * N, *storage.obs.shape[2:])
            code = self.obs_code(obs_flat)
            self.arena.bind("obs_st", obs_flat.view(-1))
        tag = tag or "t%d_" % mbs
        kw = {}
        if s.recurrent:
            # rows of the minibatch are time-major (t * E + e); masks come from the rollout buffer through
            # the row list, the initial state from `hxs_mb` (filled per minibatch by the caller)
            assert seq is not None and seq[0] * seq[1] == mbs
            self.arena.bind("masks_st", storage.masks.view(-1))
            if use_idx:
                self.arena.ensure("hxs_mb", seq[1] * s.hidden)
            else:
                self.arena.bind("hxs_mb", storage.recurrent_hidden_states[0].contiguous().view(-1))
            kw = dict(seq=seq, masks=("masks_st", 0), hxs=("hxs_mb", 0))
        nb = NetBuilder(s, self.L, mbs, "obs_st", code, use_idx, tag, full_rows=T * N if use_idx else 0, ring=ring, **kw)
        fwd = nb.forward_plans()
        bwd = nb.backward_plans() if with_backward else []
        if nb.tcx is not None:
            fwd, bwd = _tcx.hoist_preps([fwd, bwd], tag)      # all weight images: one launch at the top of the step
        if not with_backward:
            nb.buf("dz", mbs * self.zs)
        self._alloc(nb)
        prog = Program(self.dev)
        for p in fwd:
            prog.add_plan(p, self.arena)
        acts = storage.actions.view(T * N, -1)
        lo = self.arena.ensure(tag + "loss_out", 4)
        d = self._loss_desc(mbs, tag, mode, acts,
                            old_logp=storage.action_log_probs.data_ptr(),
                            old_value=storage.value_preds.data_ptr(),
                            returns=storage.returns.data_ptr(),
                            dz=self.arena.ptr((tag + "dz", 0)),
                            loss_out=lo.data_ptr(),
                            clip=float(hyper.get("clip_param", 0.0)), c_v=float(hyper["value_loss_coef"]),
                            c_e=float(hyper["entropy_coef"]),
                            use_clipped_value_loss=int(bool(hyper.get("use_clipped_value_loss", True))))
        d.idx = SENTINEL if use_idx else None
        if adv is not None:
            d.adv = adv.data_ptr()
        if noise is not None:
            d.noise = noise.data_ptr()
        if loss_accum is not None:
            d.loss_accum = loss_accum.data_ptr()
        if "inv_B" in hyper:
            d.inv_B = float(hyper["inv_B"])
        if s.action_kind == "box":
            d.dlogstd = self.arena.ptr(("G", self.L.entries["dist.logstd._bias"][0]))
        prog.add(OP_LOSS, d, RUNTIME_IDX if use_idx else 0)
        for p in bwd:
            prog.add_plan(p, self.arena)
        prog.loss_desc = d
        prog.loss_out = lo
        prog.builder = nb
        # once-per-update work (frame re-layout of the whole rollout): the caller runs it before the first minibatch
        prog.prologue = None
        if nb.tcx is not None and nb.tcx.prologue:
            prog.prologue = Program(self.dev)
            for p in nb.tcx.prologue:
                prog.prologue.add_plan(p, self.arena)
            prog.prologue.finalize()
        prog.prologue_chunk = None
        if nb.tcx is not None and nb.tcx.prologue_chunk:
            prog.prologue_chunk = Program(self.dev)
            for p in nb.tcx.prologue_chunk:
                prog.prologue_chunk.add_plan(p, self.arena)
            prog.prologue_chunk.finalize()
        return prog

    def stage_ring_batch(self, storage):
        """Frame-ring rollout read by a full-batch program: materialises observation slots [0, T) as stacked uint8 rows
        in the arena buffer the program was built against (same address every update)."""
        flat = self.arena.stage("obs_ring_flat", storage.obs[:-1])
        self.arena.bind("obs_st", flat)
        return flat

    def acktr_programs(self, storage, hyper, noise, inv_B=None):
        """ACKTR's three passes over the FULL batch (reference a2c_acktr.py:38-72) as separate programs:
        forward;  Fisher-sample loss + data-gradient-only backward (feeds the g-statistics);
        main A2C loss + full backward.  Frames are read from the fp32/255 copy `obs_f32`."""
        s = self.spec
        T, N = storage.rewards.shape[0], storage.rewards.shape[1]
        B = T * N
        tag = "t%d_" % B
        nb = NetBuilder(s, self.L, B, "obs_f32", 0, False, tag, allow_mlp_fused=False)
        fwd = nb.forward_plans()
        bwd = nb.backward_plans()
        self._alloc(nb)
        acts = storage.actions.view(B, -1)
        lo = self.arena.ensure(tag + "loss_out", 4)
        lo_f = self.arena.ensure(tag + "loss_out_fisher", 4)
        ls_rows = self.arena.ensure("kfac_dlogstd_rows", B * s.num_actions) if s.action_kind == "box" else None

        def loss(mode, out):
            d = self._loss_desc(B, tag, mode, acts, returns=storage.returns.data_ptr(),
                                dz=self.arena.ptr((tag + "dz", 0)), loss_out=out.data_ptr(),
                                c_v=float(hyper["value_loss_coef"]), c_e=float(hyper["entropy_coef"]),
                                noise=noise.data_ptr())
            if inv_B is not None:
                d.inv_B = float(inv_B)          # sharded: loss terms are local sums over the GLOBAL batch size
            if s.action_kind == "box":
                d.dlogstd = self.arena.ptr(("G", self.L.entries["dist.logstd._bias"][0]))
                if mode == 2:                   # Fisher pass: per-row gradient of the logstd AddBias (its g-factor)
                    d.dlogstd_rows = ls_rows.data_ptr()
                    d.dlogstd = None
            return d
        progs = {}
        progs["fwd"] = Program(self.dev)
        for p in fwd:
            progs["fwd"].add_plan(p, self.arena)
        progs["fisher"] = Program(self.dev)
        progs["fisher"].add(OP_LOSS, loss(2, lo_f))
        for p in bwd:
            if not (p.name.endswith(".wgrad") or ".bias_" in p.name):     # data gradients only
                progs["fisher"].add_plan(p, self.arena)
        progs["main"] = Program(self.dev)
        progs["main"].add(OP_LOSS, loss(1, lo))
        for p in bwd:
            progs["main"].add_plan(p, self.arena)
        for p in progs.values():
            p.finalize()
        progs["loss_out"] = lo
        return progs

    def add_optimizer_ops(self, prog, kind, state1, state2, max_norm):
        n = self.L.total
        scratch = self.arena.ensure("norm_scratch", 2 * (148 + 2), torch.float64)
        norm = self.arena.ensure("grad_norm", 4)
        clip = max_norm is not None and max_norm > 0
        if clip:
            prog.add(OP_GRAD_NORM, GradNormDesc(self.arena.bufs["G"].data_ptr(), n, scratch.data_ptr(), norm.data_ptr()))
        o = OptimDesc()
        o.params, o.grads = self.arena.bufs["P"].data_ptr(), self.arena.bufs["G"].data_ptr()
        o.state1 = state1.data_ptr()
        o.state2 = state2.data_ptr() if state2 is not None else None
        o.n = n
        o.norm = norm.data_ptr()
        o.kind = kind
        o.max_norm = float(max_norm) if clip else 0.0
        prog.add(OP_OPTIM, o)
        prog.optim_desc = o
        return o


from . import _tcx  # noqa: E402,F401  (registers the a2cb_tcx_* entry points; imports this module lazily)
