"""KFAC optimiser of ACKTR on the sm_100a engine.

Native counterpart of `a2c_ppo_acktr.algo.kfac.KFACOptimizer` (reference algo/kfac.py:87-242) for
the CNN policy (the reference's ACKTR recipe, README "ACKTR") and the MLP / DiagGaussian policy
(recurrent policies are excluded by the reference itself, arguments.py:156-159).  Same hyper-parameters and the same arithmetic:

  * a-factors  aa = cov(inputs)   (compute_cov_a, kfac.py:29-46): for a conv layer the patch
    covariance  sum_rows p p^T / (B (oh ow)^2), for a Linear layer x^T x / B, for a bias [[1]];
  * g-factors  gg = cov(output gradients of the Fisher-sample loss) (compute_cov_g, kfac.py:49-64):
    conv  (B oh ow) sum_rows g g^T,  Linear  B g^T g,  conv bias  B s^T s with s = sum over pixels;
  * running averages m <- decay m + (1 - decay) new (kfac.py:67-71; m = new on the first update);
  * every Tf steps an eigendecomposition of every factor (library call), eigenvalues <= 1e-6 zeroed;
  * preconditioning v = Q_g (Q_g^T dW Q_a / (d_g d_a^T + damping)) Q_a^T, trust-region factor
    nu = min(1, sqrt(kl_clip / sum(v dW lr^2))), then SGD(lr * (1 - momentum), momentum) on nu v.

Unlike the reference there is no module surgery (SplitBias) and no hooks: the layer inputs and output
gradients are the engine's own activation / gradient buffers, the patch matrix is never materialised
(implicit im2col inside the factor GEMMs) and the running average is folded into the split-K fold.
"""
import ctypes
import math

import torch

from .. import _engine, _native
from .. import _plans as P


class _Module(object):
    """One KFAC 'module' = one parameter tensor (a weight or a bias) with its two factors."""

    def __init__(self, key, rows, cols, aa, gg):
        self.key, self.rows, self.cols = key, rows, cols      # gradient viewed as [rows, cols]
        self.aa, self.gg = aa, gg                              # factor names (shared between modules)


class KFACOptimizer(object):
    _linalg_loaded = False

    def __init__(self, model, lr=0.25, momentum=0.9, stat_decay=0.99, kl_clip=0.001, damping=1e-2,
                 weight_decay=0, fast_cnn=False, Ts=1, Tf=10):
        if model.is_recurrent:
            raise _native.NativeError("ACKTR does not support recurrent policies (reference arguments.py:156-159)")
        if fast_cnn or weight_decay:
            raise _native.NativeError("fast_cnn / weight_decay are not supported (reference defaults are off)")
        self.model = model
        self.lr, self.momentum, self.stat_decay = lr, momentum, stat_decay
        self.kl_clip, self.damping, self.weight_decay = kl_clip, damping, weight_decay
        self.Ts, self.Tf = Ts, Tf
        self.steps = 0
        self.acc_stats = False
        # reference quirk kept: update_linear_schedule writes param_groups of THIS object, but the
        # stepping optimiser is the inner SGD with its own lr (kfac.py:139-142,241) -> no effect
        self.param_groups = [{"lr": lr, "params": list(model.parameters())}]
        self.factors = {}          # name -> dict(m=, Q=, d=, n=)
        self.modules = []
        self._built_for = None
        self.momentum_buf = None

    # ------------------------------------------------------------------
    def _factor(self, name, n, dev):
        if name not in self.factors:
            self.factors[name] = dict(n=n, m=torch.zeros(n * n, device=dev), Q=torch.zeros(n * n, device=dev),
                                      d=torch.zeros(n, device=dev))
        return self.factors[name]

    def _build(self, rt, rollouts, world=1):
        """Programs for the factor statistics and the preconditioning chain (built once per buffer set).
        world > 1: this rank holds 1/world of the batch; factor scales use the GLOBAL batch size and the
        running-average weight of the old value is divided by world, so that the SUM over ranks of the
        locally folded factors is decay * m + (1 - decay) * new (the caller all-reduces them)."""
        spec, L, dev = rt.spec, rt.L, rt.dev
        T, N = rollouts.rewards.shape[0], rollouts.rewards.shape[1]
        B = T * N
        Bg = B * world
        self.world = world
        key = (id(rt), rollouts.obs.data_ptr(), B, world)
        if self._built_for == key:
            return
        A, H = spec.num_actions, spec.hidden
        tag = "t%d_" % B
        nb = _engine.NetBuilder(spec, L, B, "obs_f32", 0, False, tag, allow_mlp_fused=False)
        nb.forward_plans()
        nb.backward_plans()
        arena = rt.arena
        zs = 1 + A

        def buf(name, numel):
            return arena.ensure("kfac_" + name, numel)

        self.modules = []
        a_plans, g_plans, self._spatial = [], [], []
        one = self._factor("one", 1, dev)
        one["m"].fill_(1.0)

        def factor_plan(name, n, Asrc, Bsrc, rowA, redA, redB, colB, rowC, colC, K, scale, vecA, vecB):
            f = self._factor(name, n, dev)
            arena.bind("kfac_m_" + name, f["m"])
            ktiles = (K + 15) // 16
            tiles = ((n + 127) // 128) * ((n + 127) // 128)
            splits = int(max(2, min((2 * 148 + tiles - 1) // tiles, max(2, ktiles // 8), 64)))
            part = buf("part_" + name, splits * n * n)
            p = P.Plan("kfac." + name, Asrc, Bsrc, ("kfac_part_" + name, 0), n, n, K, rowA, redA, redB, colB, rowC, colC,
                       splits=splits, split_stride=n * n, split_dst=("kfac_m_" + name, 0, n * n), accumulate=1,
                       vecA=vecA, vecB=vecB, alpha=scale, tile=1001, sym=1)
            p.base_scale = scale
            return p

        if not spec.cnn:
            self._mlp_modules(spec, tag, B, Bg, factor_plan, a_plans, g_plans)
        c1, c2, c3 = (nb.c1, nb.c2, nb.c3) if spec.cnn else (None, None, None)
        layers = [] if not spec.cnn else [
                  ("conv1", c1, ("obs_f32", 0), ("base.main.0.weight", "base.main.0.bias")),
                  ("conv2", c2, (tag + "a1", 0), ("base.main.2.weight", "base.main.2.bias")),
                  ("conv3", c3, (tag + "a2", 0), ("base.main.4.weight", "base.main.4.bias"))]
        gbufs = {"conv1": (tag + "g1", 0), "conv2": (tag + "g2", 0), "conv3": (tag + "g3", 0)}
        for name, cv, src, (wk, bk) in layers:
            Pn = cv.oh * cv.ow
            k, c = cv.k, cv.cin
            if cv.in_layout == "nchw":
                rowC, colC = P.plain(cv.K), P.plain(1)
            else:   # engine K order (ky, kx, c) -> reference patch order (c, ky, kx)
                rowC, colC = P.aff(k * c, c, k * cv.K, cv.K, k * k * cv.K), P.aff(k * c, c, k, 1, k * k)
            v = cv.vec_ok()
            a_plans.append(factor_plan("aa_" + name, cv.K, src, src, cv.red_in(), cv.row_in(), cv.row_in(), cv.red_in(),
                                       rowC, colC, B * Pn, 1.0 / (Bg * float(Pn) ** 2), 2 if v else 0, 1 if v else 0))
            g = (gbufs[name][0], gbufs[name][1] + cv.gout_base)
            g_plans.append(factor_plan("gg_" + name, cv.cout, g, g, P.plain(1), cv.row_gout(), cv.row_gout(), P.plain(1),
                                       P.plain(cv.cout), P.plain(1), B * Pn, float(Bg * Pn), 2, 1))
            s = buf("s_" + name, B * cv.cout)
            self._spatial.append((s, gbufs[name][0], cv))
            g_plans.append(factor_plan("ggb_" + name, cv.cout, ("kfac_s_" + name, 0), ("kfac_s_" + name, 0), P.plain(1),
                                       P.plain(cv.cout), P.plain(cv.cout), P.plain(1), P.plain(cv.cout), P.plain(1), B,
                                       float(Bg), 2, 1))
            self.modules.append(_Module(wk, cv.cout, cv.K, "aa_" + name, "gg_" + name))
            self.modules.append(_Module(bk, cv.cout, 1, "one", "ggb_" + name))
        if spec.cnn:
            self._cnn_dense_modules(spec, tag, B, Bg, factor_plan, a_plans, g_plans)

        self._a_plans, self._g_plans = a_plans, g_plans
        self._finish_build(rt, key, a_plans, g_plans, buf)

    def _mlp_modules(self, spec, tag, B, Bg, factor_plan, a_plans, g_plans):
        """MLPBase + DiagGaussian (reference model.py:198-229, distributions.py:76-95): six Linear layers, their
        biases, and the logstd AddBias -- 13 modules.  Inputs and output gradients are the
        layer-by-layer engine's buffers; the two first layers share their input statistics."""
        A, H, d = spec.num_actions, spec.hidden, spec.obs_shape[0]
        zs = 1 + A
        pl = P.plain

        def aa(name, src, n):
            v = (2, 1) if n % 4 == 0 else (0, 0)
            a_plans.append(factor_plan(name, n, src, src, pl(1), pl(n), pl(n), pl(1), pl(n), pl(1), B, 1.0 / Bg, *v))

        def gg(name, src, n, stride):
            v = (2, 1) if (n % 4 == 0 and stride == n) else (0, 0)
            g_plans.append(factor_plan(name, n, src, src, pl(1), pl(stride), pl(stride), pl(1), pl(n), pl(1), B, float(Bg), *v))
        aa("aa_x", ("obs_f32", 0), d)
        for nm in ("ha1", "hc1", "ha2", "hc2"):
            aa("aa_" + nm, (tag + nm, 0), H)
        for nm in ("ga1", "ga2", "gc1", "gc2"):
            gg("gg_" + nm, (tag + nm, 0), H, H)
        gg("gg_critic", (tag + "dz", 0), 1, zs)
        gg("gg_dist", (tag + "dz", 1), A, zs)
        gg("gg_logstd", ("kfac_dlogstd_rows", 0), A, A)
        for key, r, c, fa, fg in (("base.actor.0", H, d, "aa_x", "gg_ga1"), ("base.actor.2", H, H, "aa_ha1", "gg_ga2"),
                                  ("base.critic.0", H, d, "aa_x", "gg_gc1"), ("base.critic.2", H, H, "aa_hc1", "gg_gc2"),
                                  ("base.critic_linear", 1, H, "aa_hc2", "gg_critic"),
                                  ("dist.fc_mean", A, H, "aa_ha2", "gg_dist")):
            self.modules.append(_Module(key + ".weight", r, c, fa, fg))
            self.modules.append(_Module(key + ".bias", r, 1, "one", fg))
        self.modules.append(_Module("dist.logstd._bias", A, 1, "one", "gg_logstd"))

    def _cnn_dense_modules(self, spec, tag, B, Bg, factor_plan, a_plans, g_plans):
        A, H = spec.num_actions, spec.hidden
        zs = 1 + A
        # fc: input a3 (NCHW-flatten order == the reference's), output gradient gh
        a_plans.append(factor_plan("aa_fc", 1568, (tag + "a3", 0), (tag + "a3", 0), P.plain(1), P.plain(1568),
                                   P.plain(1568), P.plain(1), P.plain(1568), P.plain(1), B, 1.0 / Bg, 2, 1))
        g_plans.append(factor_plan("gg_fc", H, (tag + "gh", 0), (tag + "gh", 0), P.plain(1), P.plain(H), P.plain(H),
                                   P.plain(1), P.plain(H), P.plain(1), B, float(Bg), 2, 1))
        self.modules.append(_Module("base.main.7.weight", H, 1568, "aa_fc", "gg_fc"))
        self.modules.append(_Module("base.main.7.bias", H, 1, "one", "gg_fc"))
        # heads share their input h
        a_plans.append(factor_plan("aa_heads", H, (tag + "h", 0), (tag + "h", 0), P.plain(1), P.plain(H), P.plain(H),
                                   P.plain(1), P.plain(H), P.plain(1), B, 1.0 / Bg, 2, 1))
        g_plans.append(factor_plan("gg_critic", 1, (tag + "dz", 0), (tag + "dz", 0), P.plain(1), P.plain(zs), P.plain(zs),
                                   P.plain(1), P.plain(1), P.plain(1), B, float(Bg), 0, 0))
        g_plans.append(factor_plan("gg_dist", A, (tag + "dz", 1), (tag + "dz", 1), P.plain(1), P.plain(zs), P.plain(zs),
                                   P.plain(1), P.plain(A), P.plain(1), B, float(Bg), 0, 0))
        hw = "dist.linear"
        self.modules.append(_Module("base.critic_linear.weight", 1, H, "aa_heads", "gg_critic"))
        self.modules.append(_Module("base.critic_linear.bias", 1, 1, "one", "gg_critic"))
        self.modules.append(_Module(hw + ".weight", A, H, "aa_heads", "gg_dist"))
        self.modules.append(_Module(hw + ".bias", A, 1, "one", "gg_dist"))

    def _finish_build(self, rt, key, a_plans, g_plans, buf):
        L, dev, arena = rt.L, rt.dev, rt.arena
        self._a_prog, self._g_prog = _engine.Program(dev), _engine.Program(dev)
        for p in a_plans:
            self._a_prog.add_plan(p, arena)
        for p in g_plans:
            self._g_prog.add_plan(p, arena)
        self._a_prog.finalize()
        self._g_prog.finalize()

        # ---- preconditioning chain per module (kfac.py:216-227) ----
        V = arena.ensure("kfac_V", L.total)
        self.V = V
        self._pre = []
        pre = _engine.Program(dev)
        for m in self.modules:
            off = L.entries[m.key][0]
            fa, fg = self.factors[m.aa], self.factors[m.gg]
            arena.bind("kfac_Qa_" + m.aa, fa["Q"])
            arena.bind("kfac_Qg_" + m.gg, fg["Q"])
            r, c = m.rows, m.cols
            t1, t2 = buf("t1", 512 * 1568), buf("t2", 512 * 1568)
            G, Qa, Qg = ("G", off), ("kfac_Qa_" + m.aa, 0), ("kfac_Qg_" + m.gg, 0)
            T1, T2, Vr = ("kfac_t1", 0), ("kfac_t2", 0), ("kfac_V", off)
            pl = P.plain
            if c > 1:
                # t1 = dW Q_a ; t2 = Q_g^T t1 ; (scale) ; t1 = t2 Q_a^T ; v = Q_g t1
                pre.add_plan(P.Plan("pre.gQa", G, Qa, T1, r, c, c, pl(c), pl(1), pl(c), pl(1), pl(c), pl(1),
                                    vecA=1 if c % 4 == 0 else 0, vecB=1 if c % 4 == 0 else 0, tile=1001), arena)
            src = T1 if c > 1 else G
            pre.add_plan(P.Plan("pre.QgT", Qg, src, T2, r, c, r, pl(1), pl(r), pl(c), pl(1), pl(c), pl(1), tile=1001), arena)
            self._pre.append(("scale", T2, fg["d"], fa["d"] if c > 1 else None, r, c))
            pre_scale_index = len(pre.entries)
            pre.add(_engine.OP_FILL, _engine.FillDesc(None, 0, 0, 0))        # placeholder slot, replaced below
            if c > 1:
                pre.add_plan(P.Plan("pre.QaT", T2, Qa, T1, r, c, c, pl(c), pl(1), pl(1), pl(c), pl(c), pl(1),
                                    vecA=1 if c % 4 == 0 else 0, vecB=2 if c % 4 == 0 else 0, tile=1001), arena)
            src2 = T1 if c > 1 else T2
            pre.add_plan(P.Plan("pre.Qg", Qg, src2, Vr, r, c, r, pl(r), pl(1), pl(c), pl(1), pl(c), pl(1), tile=1001), arena)
            self._pre[-1] = self._pre[-1] + (pre_scale_index,)
        self._pre_prog = pre
        self._built_for = key

    # ------------------------------------------------------------------
    def _run_pre(self, rt):
        """Runs the preconditioning chain; the element-wise divide sits between GEMMs of each module."""
        lib, st = _native.lib(), _native.stream_ptr(rt.dev)
        entries = self._pre_prog.entries
        # split the op list at the scale placeholders
        start = 0
        for (_, T2, dg, da, r, c, idx) in self._pre:
            seg = _engine.Program(rt.dev)
            seg.entries = entries[start:idx]
            seg.run()
            _native.check(lib.a2cb_kfac_scale(rt.arena.ptr(T2), dg.data_ptr(), da.data_ptr() if da is not None else None,
                                              r, c, float(self.damping + self.weight_decay), st), "a2cb_kfac_scale")
            start = idx + 1
        seg = _engine.Program(rt.dev)
        seg.entries = entries[start:]
        seg.run()

    def accumulate_a(self, rt):
        """a-statistics of the forward pass just run (reference _save_input, kfac.py:144-159)."""
        if self.steps % self.Ts != 0:
            return
        self._set_weights(self._a_prog, self._a_plans)
        self._a_prog.run()

    def accumulate_g(self, rt):
        """g-statistics of the Fisher backward just run (reference _save_grad_output, kfac.py:161-177)."""
        lib, st = _native.lib(), _native.stream_ptr(rt.dev)
        for s, gname, cv in self._spatial:
            g = rt.arena.bufs[gname]
            _native.check(lib.a2cb_spatial_sum(s.data_ptr(), g.data_ptr() + 4 * cv.gout_base, s.numel() // cv.cout,
                                               cv.oh * cv.ow, cv.cout, cv.ow, cv.gout_row, cv.pw * cv.cout, st),
                          "a2cb_spatial_sum")
        self._set_weights(self._g_prog, self._g_plans)
        self._g_prog.run()

    def _set_weights(self, prog, plans):
        """Running average m <- decay m + (1 - decay) new; m = new on the very first update."""
        first = self.steps == 0
        w_new = 1.0 if first else (1.0 - self.stat_decay)
        beta = 0.0 if first else self.stat_decay / getattr(self, "world", 1)
        gi = 0
        for kind, flags, desc in prog.entries:
            if kind == _engine.OP_GEMM:
                desc.alpha = plans[gi].base_scale * w_new
                gi += 1
            elif kind == _engine.OP_REDUCE:
                desc.beta = beta

    def step(self, rt):
        """Preconditions the flat gradient and applies the SGD-momentum step (kfac.py:190-242)."""
        L = rt.L
        if self.steps % self.Tf == 0:
            self._eigendecompose(rt.dev)
        self._run_pre(rt)
        lib, st = _native.lib(), _native.stream_ptr(rt.dev)
        scratch = rt.arena.ensure("kfac_nu_scratch", 160, torch.float64)
        nu = rt.arena.ensure("kfac_nu", 4)
        G = rt.arena.bufs["G"]
        _native.check(lib.a2cb_kfac_nu(self.V.data_ptr(), G.data_ptr(), L.total, scratch.data_ptr(), nu.data_ptr(),
                                       float(self.lr), float(self.kl_clip), st), "a2cb_kfac_nu")
        if self.momentum_buf is None:
            self.momentum_buf = torch.zeros_like(G)
        first = int(self.steps == 0)
        lr_in = self.lr * (1 - self.momentum)
        _native.check(lib.a2cb_optim_step(2, rt.arena.bufs["P"].data_ptr(), self.V.data_ptr(),
                                          self.momentum_buf.data_ptr(), None, L.total, nu.data_ptr(), lr_in, 0.0,
                                          self.momentum, 0.0, 1.0 - self.momentum, 1.0, 1.0, -1.0, first, st),
                      "a2cb_optim_step")
        G.copy_(self.V)          # p.grad <- nu * v, as the reference leaves it (kfac.py:236-239)
        self.steps += 1

    def _eigendecompose(self, dev):
        """Eigendecomposition of every factor (reference kfac.py:208-211) -- the one library call KFAC keeps
        (`torch.linalg.eigh`, cuSOLVER).  The 12-13 decompositions are independent, latency-bound chains of small
        kernels (69 ms back to back on B200 for the CNN policy: 19 ms for the 1569 x 1569 factor alone,
        tools/eigh_bench.py), and every call ends with a host-side status check, so they run on worker threads with one
        CUDA stream each, largest factor first; the update's stream waits for all of them."""
        import concurrent.futures
        cur = torch.cuda.current_stream(dev)
        names = sorted(self.factors, key=lambda k: -self.factors[k]["n"])
        nw = min(6, len(names))
        if not KFACOptimizer._linalg_loaded:
            torch.linalg.eigh(torch.eye(2, device=dev))      # torch loads its linalg library lazily, and not thread-safely
            KFACOptimizer._linalg_loaded = True
        if getattr(self, "_eigh_pool", None) is None:
            self._eigh_pool = concurrent.futures.ThreadPoolExecutor(max_workers=nw)
            self._eigh_streams = [torch.cuda.Stream(device=dev) for _ in range(nw)]
        ready = torch.cuda.Event()
        ready.record(cur)

        def work(slot):
            s = self._eigh_streams[slot]
            s.wait_event(ready)
            with torch.cuda.device(dev), torch.cuda.stream(s):
                for name in names[slot::nw]:
                    f = self.factors[name]
                    n = f["n"]
                    d, Q = torch.linalg.eigh(f["m"].view(n, n), UPLO="U")
                    d = d * (d > 1e-6).float()
                    f["d"].copy_(d)
                    f["Q"].copy_(Q.contiguous().view(-1))
            return slot
        for fut in [self._eigh_pool.submit(work, i) for i in range(nw)]:
            fut.result()
        for s in self._eigh_streams:
            cur.wait_stream(s)

    def __getstate__(self):
        st = dict(self.__dict__)
        st.pop("_eigh_pool", None)
        st.pop("_eigh_streams", None)
        return st

    def zero_grad(self, set_to_none=False):
        self.model.runtime().arena.bufs["G"].zero_()
