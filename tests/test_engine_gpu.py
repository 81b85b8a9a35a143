"""GPU: the affine-indexed GEMM engine, the fused loss epilogue and the optimiser kernels
(all through the C ABI) against the numpy emulator / torch fp32 references."""
import ctypes
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from a2c_ppo_acktr import _engine, _native  # noqa: E402
from a2c_ppo_acktr import _plans as P  # noqa: E402

DEV = "cuda:0"


@pytest.fixture(params=["tcgen05", "simt"], autouse=True)
def gemm_engine(request):
    """Every test of this file runs on both engines behind a2cb_gemm."""
    _native.set_gemm_mode(1 if request.param == "tcgen05" else 0)
    yield request.param
    _native.set_gemm_mode(2)


def run_plan(plan, bufs, idx=None):
    """Runs one plan on the GPU on copies of the numpy buffers; returns the updated buffers."""
    arena = _engine.Arena(torch.device(DEV))
    for k, v in bufs.items():
        arena.bind(k, torch.from_numpy(v.copy()).to(DEV))
    prog = _engine.Program(torch.device(DEV))
    prog.add_plan(plan, arena)
    prog.run(torch.from_numpy(np.asarray(idx, dtype=np.int64)).to(DEV) if idx is not None else None)
    torch.cuda.synchronize()
    return {k: arena.bufs[k].cpu().numpy() for k in bufs}


def rand_affine(rng, n, max_off):
    """A random injective-enough 3-digit map for n indices."""
    d2 = int(rng.integers(1, 5))
    d1 = d2 * int(rng.integers(1, 4))
    s2 = int(rng.integers(1, 4))
    s1 = s2 * d2 + int(rng.integers(0, 3))
    s0 = s1 * (d1 // d2) + int(rng.integers(0, 5))
    return (d1, d2, s0, s1, s2), ((n - 1) // d1) * s0 + ((d1 - 1) // d2) * s1 + (d2 - 1) * s2 + 1


@pytest.mark.parametrize("M,N,K,tile", [(1, 1, 1, 0), (129, 7, 33, 0), (300, 33, 70, 64), (257, 130, 19, 128),
                                        (64, 64, 64, 32), (500, 96, 128, 0)])
@pytest.mark.parametrize("variant", ["plain", "epilogue", "ones", "batch"])
def test_gemm_random_maps_vs_emulator(M, N, K, tile, variant, gemm_engine):
    rng = np.random.default_rng(M * 1000 + N * 10 + K)
    rowA, spanRA = rand_affine(rng, M, 0)
    redA, spanKA = rand_affine(rng, K, 0)
    redB, spanKB = rand_affine(rng, K, 0)
    colB, spanCB = rand_affine(rng, N, 0)
    # output map must be injective: plain row-major with padding
    ldc = N + int(rng.integers(0, 3))
    rowC, colC = P.plain(ldc), P.plain(1)
    nb = 3 if variant == "batch" else 1
    bufs = {
        "A": rng.standard_normal(spanRA + spanKA + 64 * nb).astype(np.float32),
        "B": rng.standard_normal(spanKB + spanCB + 64 * nb).astype(np.float32),
        "C": np.zeros(M * ldc * nb + 8, np.float32),
        "bias": rng.standard_normal(max(M, N)).astype(np.float32),
        "mask": rng.standard_normal(M * ldc).astype(np.float32),
        "ones": np.zeros(N * 2, np.float32),
    }
    kw = dict(tile=tile)
    if variant == "epilogue":
        kw.update(bias=("bias", 0), bias_mode=1 + (M % 2), act=P.ACT_RELU if N % 2 else P.ACT_TANH,
                  mask=("mask", 0), mask_mode=1 + (K % 2), rowM=rowC, colM=colC, alpha=0.5)
    if variant == "ones":
        kw.update(ones=("ones", 1), ones_stride=2 if N > 1 else 1)
    if variant == "batch":
        kw.update(batch=3, batch_offA=[0, 17, 40, 0], batch_offB=[0, 5, 64, 0], batch_offC=[0, M * ldc, 2 * M * ldc, 0])
    plan = P.Plan("t", ("A", 0), ("B", 0), ("C", 0), M, N, K, rowA, redA, redB, colB, rowC, colC, **kw)
    want = {k: v.copy() for k, v in bufs.items()}
    P.emulate(plan, want)
    got = run_plan(plan, bufs)
    # 3xTF32 drops the lo*lo term and rounds lo to 11 bits: ~2^-21 per product, sums of K = 128 N(0,1) products
    atol = 1e-4 if gemm_engine == "tcgen05" else 2e-5
    np.testing.assert_allclose(got["C"], want["C"], rtol=2e-5, atol=atol)
    np.testing.assert_allclose(got["ones"], want["ones"], rtol=2e-5, atol=atol)


def test_gemm_split_k_and_reduce():
    rng = np.random.default_rng(5)
    M, N, K, splits = 70, 33, 1000, 7
    span = M * N + N
    bufs = {"A": rng.standard_normal(M * K).astype(np.float32), "B": rng.standard_normal(K * N).astype(np.float32),
            "part": np.zeros(splits * span, np.float32), "G": np.full(span + 5, 3.0, np.float32)}
    plan = P.Plan("t", ("A", 0), ("B", 0), ("part", 0), M, N, K, P.plain(K), P.plain(1), P.plain(N), P.plain(1),
                  P.plain(N), P.plain(1), ones=("part", M * N), splits=splits, split_stride=span,
                  split_dst=("G", 5, span))
    got = run_plan(plan, bufs)
    A = bufs["A"].reshape(M, K).astype(np.float64)
    B = bufs["B"].reshape(K, N).astype(np.float64)
    np.testing.assert_allclose(got["G"][5:5 + M * N].reshape(M, N), A @ B, rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(got["G"][5 + M * N:], B.sum(0), rtol=1e-5, atol=1e-4)
    assert np.all(got["G"][:5] == 3.0)


@pytest.mark.parametrize("n,K,perm", [(1, 300, False), (96, 700, False), (257, 3000, True), (576, 2100, True)])
def test_gemm_symmetric_product_upper_tiles_and_mirror(n, K, perm):
    """KFAC factor products (reference kfac.py:29-64: a^T a, g^T g) as SYMMETRIC products: `sym` makes the engine skip the
    tiles strictly below the diagonal and lets the tiles above store their mirror image through the transposed output maps
    (a2cb_gemm_t.sym).  The folded result must equal the full product -- here with split-K, a running average folded in
    (beta * old + sum) and, for `perm`, the digit-reversing output permutation the conv factors use (engine K order
    (ky, kx, c) -> reference patch order (c, ky, kx))."""
    rng = np.random.default_rng(n + K)
    splits = 3 if K > 100 else 1
    X = rng.standard_normal((K, n)).astype(np.float32)
    old = rng.standard_normal(n * n).astype(np.float32)
    old = (old.reshape(n, n) + old.reshape(n, n).T).reshape(-1).copy()
    bufs = {"X": X.reshape(-1).copy(), "part": np.full(splits * n * n, 7.0, np.float32), "m": old.copy()}
    if perm and n % 9 == 0:         # n = c * 9: (tap, c) -> (c, tap)
        c = n // 9
        rowC, colC = P.aff(c, 1, n, 9 * n, 0), P.aff(c, 1, 1, 9, 0)
        order = np.array([(i % c) * 9 + i // c for i in range(n)])
    else:
        rowC, colC = P.plain(n), P.plain(1)
        order = np.arange(n)
    plan = P.Plan("sym", ("X", 0), ("X", 0), ("part", 0), n, n, K, P.plain(1), P.plain(n), P.plain(n), P.plain(1), rowC, colC,
                  splits=splits, split_stride=n * n, split_dst=("m", 0, n * n), accumulate=1, split_beta=0.5, alpha=0.25,
                  tile=1001, sym=1, vecA=2 if n % 4 == 0 else 0, vecB=1 if n % 4 == 0 else 0)
    got = run_plan(plan, bufs)
    full = 0.25 * (X.astype(np.float64).T @ X.astype(np.float64))
    want = np.zeros((n, n))
    want[np.ix_(order, order)] = full
    want = 0.5 * old.reshape(n, n) + want
    np.testing.assert_allclose(got["m"].reshape(n, n), want, rtol=2e-5, atol=2e-4)
    assert np.array_equal(got["m"].reshape(n, n), got["m"].reshape(n, n).T), "the mirror image is a copy: exactly symmetric"


def test_gru_tc_consecutive_launches_with_changing_group_counts():
    """The tcgen05 recurrence validates every exchanged word by a 2-bit step tag and relies on a memset of the exchange
    buffers in front of each launch (csrc/gru_tc.cu): launches with more, fewer and again more environment groups, odd
    and even step counts, must not see each other's leftovers."""
    _native.set_gru_mode(2)
    try:
        for T, N, H in ((6, 150, 512), (5, 8, 512), (7, 64, 512), (3, 150, 512), (2, 24, 256), (9, 64, 512)):
            if _native.gru_seq_variant(N, H) != "tc":
                pytest.skip("tcgen05 recurrence not available for this shape")
            _run_gru_case(T, N, H, 0.2)
    finally:
        _native.set_gru_mode(-1)


@pytest.mark.parametrize("B", [4, 37])
def test_cnn_forward_backward_conv3_patch_mode(B, monkeypatch):
    """The opt-in TF32 patch mode of conv3's forward (A2CB_TCX_PATCH3=1: one 9 x 9 patch per tile and channel half,
    the nine taps as row-shifted UMMA windows, weight k-blocks through their own ring) against the same reference.
    Correct, but measured slower than the per-tap boxes (0.285 vs 0.204 ms: 40 % of a patch's accumulator rows are
    junk and the small-N layers are epilogue-bound), hence off by default."""
    monkeypatch.setenv("A2CB_TCX_PATCH3", "1")
    test_cnn_forward_backward_vs_torch(torch.uint8, B)


@pytest.mark.parametrize("B", [5, 1, 37])
@pytest.mark.parametrize("obs_dtype", [torch.uint8, torch.float32])
def test_cnn_forward_backward_vs_torch(obs_dtype, B):
    """Whole CNN trunk + heads through the engine (gathered rows, fused /255) vs torch autograd; B covers a single
    row, partial image tiles and more than one tile of every layer."""
    from a2c_ppo_acktr.model import Policy

    class Discrete(object):
        n = 6
    torch.manual_seed(1)
    pol = Policy((4, 84, 84), Discrete())
    sd = {k: v.clone() for k, v in pol.state_dict().items()}
    # non-zero biases so that the bias paths are exercised
    g = torch.Generator().manual_seed(3)
    for k in sd:
        if k.endswith("bias"):
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.1
    pol.load_state_dict(sd)
    pol.to(DEV)
    rows = 11
    obs_u8 = torch.randint(0, 256, (rows, 4, 84, 84), generator=g, dtype=torch.uint8)
    idx = torch.tensor([7, 0, 3, 10, 3] * 8)[:B]
    action = torch.randint(0, 6, (B, 1), generator=g)
    # torch reference on CPU
    ref = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    x = obs_u8[idx].float() / 255.0
    a = F.relu(F.conv2d(x, ref["base.main.0.weight"], ref["base.main.0.bias"], stride=4))
    a = F.relu(F.conv2d(a, ref["base.main.2.weight"], ref["base.main.2.bias"], stride=2))
    a = F.relu(F.conv2d(a, ref["base.main.4.weight"], ref["base.main.4.bias"], stride=1))
    h = F.relu(F.linear(a.reshape(B, -1), ref["base.main.7.weight"], ref["base.main.7.bias"]))
    value = F.linear(h, ref["base.critic_linear.weight"], ref["base.critic_linear.bias"])
    logits = F.linear(h, ref["dist.linear.weight"], ref["dist.linear.bias"])
    dist = torch.distributions.Categorical(logits=logits)
    logp = dist.log_prob(action.squeeze(-1)).unsqueeze(-1)
    ent = dist.entropy().mean()
    cv, cl, ce = torch.randn(B, 1, generator=g), torch.randn(B, 1, generator=g), 0.7
    ((value * cv).sum() + (logp * cl).sum() + ce * ent).backward()

    obs_dev = (obs_u8 if obs_dtype == torch.uint8 else obs_u8.float())[idx].to(DEV)
    v2, lp2, ent2, _ = pol.evaluate_actions(obs_dev, torch.zeros(B, 1, device=DEV), torch.ones(B, 1, device=DEV),
                                            action.to(DEV))
    tol = dict(rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(v2.detach().cpu().numpy(), value.detach().numpy(), **tol)
    np.testing.assert_allclose(lp2.detach().cpu().numpy(), logp.detach().numpy(), **tol)
    np.testing.assert_allclose(ent2.item(), ent.item(), **tol)
    for p in pol.parameters():
        p.grad.zero_()
    ((v2 * cv.to(DEV)).sum() + (lp2 * cl.to(DEV)).sum() + ce * ent2).backward()
    for k, p in pol.named_parameters():
        np.testing.assert_allclose(p.grad.cpu().numpy(), ref[k].grad.numpy(), rtol=2e-4, atol=2e-5, err_msg=k)


def test_mlp_forward_backward_vs_torch():
    from a2c_ppo_acktr.model import Policy

    class Box(object):
        shape = (3,)
    torch.manual_seed(1)
    pol = Policy((11,), Box())
    g = torch.Generator().manual_seed(4)
    sd = {k: v.clone() for k, v in pol.state_dict().items()}
    for k in sd:
        if k.endswith("bias"):
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.1
    pol.load_state_dict(sd)
    pol.to(DEV)
    B = 9
    obs = torch.randn(B, 11, generator=g)
    action = torch.randn(B, 3, generator=g)
    ref = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    hc = torch.tanh(F.linear(obs, ref["base.critic.0.weight"], ref["base.critic.0.bias"]))
    hc = torch.tanh(F.linear(hc, ref["base.critic.2.weight"], ref["base.critic.2.bias"]))
    ha = torch.tanh(F.linear(obs, ref["base.actor.0.weight"], ref["base.actor.0.bias"]))
    ha = torch.tanh(F.linear(ha, ref["base.actor.2.weight"], ref["base.actor.2.bias"]))
    value = F.linear(hc, ref["base.critic_linear.weight"], ref["base.critic_linear.bias"])
    mean = F.linear(ha, ref["dist.fc_mean.weight"], ref["dist.fc_mean.bias"])
    std = (torch.zeros_like(mean) + ref["dist.logstd._bias"].t().view(1, -1)).exp()
    dist = torch.distributions.Normal(mean, std)
    logp = dist.log_prob(action).sum(-1, keepdim=True)
    ent = dist.entropy().sum(-1).mean()
    cv, cl, ce = torch.randn(B, 1, generator=g), torch.randn(B, 1, generator=g), -0.3
    ((value * cv).sum() + (logp * cl).sum() + ce * ent).backward()
    v2, lp2, ent2, _ = pol.evaluate_actions(obs.to(DEV), torch.zeros(B, 1, device=DEV), torch.ones(B, 1, device=DEV),
                                            action.to(DEV))
    tol = dict(rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(v2.detach().cpu().numpy(), value.detach().numpy(), **tol)
    np.testing.assert_allclose(lp2.detach().cpu().numpy(), logp.detach().numpy(), **tol)
    np.testing.assert_allclose(ent2.item(), ent.item(), **tol)
    for p in pol.parameters():
        p.grad.zero_()
    ((v2 * cv.to(DEV)).sum() + (lp2 * cl.to(DEV)).sum() + ce * ent2).backward()
    for k, p in pol.named_parameters():
        np.testing.assert_allclose(p.grad.cpu().numpy(), ref[k].grad.numpy(), rtol=2e-4, atol=2e-5, err_msg=k)


@pytest.mark.parametrize("kind", ["adam", "rmsprop"])
def test_optimizer_step_vs_torch(kind):
    n = 10007
    g = torch.Generator().manual_seed(0)
    p0 = torch.randn(n, generator=g)
    ref = p0.clone().requires_grad_(True)
    opt = (torch.optim.Adam([ref], lr=2.5e-4, eps=1e-5) if kind == "adam"
           else torch.optim.RMSprop([ref], 7e-4, eps=1e-5, alpha=0.99))
    p = p0.clone().to(DEV)
    s1, s2 = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    scratch = torch.zeros(160, dtype=torch.float64, device=DEV)
    norm = torch.zeros(4, device=DEV)
    lib = _native.lib()
    for t in range(1, 4):
        grad = torch.randn(n, generator=g) * (3.0 if t == 2 else 0.001)
        ref.grad = grad.clone()
        tn = torch.nn.utils.clip_grad_norm_([ref], 0.5)
        opt.step()
        gd = grad.clone().to(DEV)
        st = _native.stream_ptr(torch.device(DEV))
        _native.check(lib.a2cb_grad_norm(gd.data_ptr(), n, scratch.data_ptr(), norm.data_ptr(), st), "norm")
        if kind == "adam":
            bc1, bc2 = 1 - 0.9 ** t, 1 - 0.999 ** t
            args = (0, 2.5e-4 / bc1, 1e-5, 0.9, 0.999, 1 - 0.9, 1 - 0.999, math.sqrt(bc2))
        else:
            args = (1, 7e-4, 1e-5, 0.0, 0.99, 1.0, 1 - 0.99, 1.0)
        _native.check(lib.a2cb_optim_step(args[0], p.data_ptr(), gd.data_ptr(), s1.data_ptr(), s2.data_ptr(), n,
                                          norm.data_ptr(), *args[1:], 0.5, int(t == 1), st), "optim")
        torch.cuda.synchronize()
        np.testing.assert_allclose(norm[0].item(), float(tn), rtol=1e-6)
        np.testing.assert_allclose(gd.cpu().numpy(), ref.grad.numpy(), rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(p.cpu().numpy(), ref.detach().numpy(), rtol=1e-6, atol=1e-7)


def test_adv_normalize_vs_torch():
    g = torch.Generator().manual_seed(2)
    for n in (7, 1024, 40001):
        ret, val = torch.randn(n, generator=g) * 3 + 1, torch.randn(n, generator=g)
        adv = ret - val
        want = (adv - adv.mean()) / (adv.std() + 1e-5)
        rd, vd = ret.to(DEV), val.to(DEV)
        out = torch.zeros(n, device=DEV)
        scratch = torch.zeros(2 * 128 + 2, dtype=torch.float64, device=DEV)
        mom = torch.zeros(4, dtype=torch.float64, device=DEV)
        lib, st = _native.lib(), _native.stream_ptr(torch.device(DEV))
        _native.check(lib.a2cb_adv_moments(rd.data_ptr(), vd.data_ptr(), n, scratch.data_ptr(), mom.data_ptr(), st), "m")
        _native.check(lib.a2cb_adv_normalize(rd.data_ptr(), vd.data_ptr(), n, mom.data_ptr(), out.data_ptr(), st), "n")
        np.testing.assert_allclose(out.cpu().numpy(), want.numpy(), rtol=1e-5, atol=1e-6)


# ---- persistent whole-sequence GRU (csrc/gru_persistent.cu) against torch autograd -------------------
@pytest.mark.parametrize("T,N,H,p_done", [(5, 3, 128, 0.2), (16, 64, 512, 0.2), (7, 77, 256, 0.2), (4, 130, 128, 0.2),
                                          (128, 64, 512, 0.01), (128, 8, 512, 0.0), (128, 150, 512, 0.01)])
@pytest.mark.parametrize("mode", ["tc", "mma", "simt"])
def test_gru_persistent_vs_torch(T, N, H, p_done, mode):
    """a2cb_gru_seq forward and backward-through-time on the masked recurrence of model.py:111-166;
    N covers partial warps, partial 64-environment blocks and more than one block.  The T = 128 cases are the
    benchmarked C5 minibatch shape (64 environments, H = 512, episodes ~100 steps long: a 128-step BPTT chain), a
    single environment group without any episode boundary, and more groups than resident CTA groups."""
    import ctypes
    from a2c_ppo_acktr import _engine as E
    _native.set_gru_mode({"tc": 2, "mma": 1, "simt": 0}[mode])
    try:
        if not E.gru_seq_supported(N, H):
            pytest.skip("cooperative launch does not fit")
        if _native.gru_seq_variant(N, H) != mode:
            pytest.skip("this shape runs the %s kernels" % _native.gru_seq_variant(N, H))
        _run_gru_case(T, N, H, p_done)
    finally:
        _native.set_gru_mode(-1)


def _run_gru_case(T, N, H, p_done):
    import ctypes
    from a2c_ppo_acktr import _engine as E
    g = torch.Generator().manual_seed(5)
    gi = torch.randn(T * N, 3 * H, generator=g)
    Whh = torch.randn(3 * H, H, generator=g) / H ** 0.5
    bhh = torch.randn(3 * H, generator=g) * 0.1
    h0 = torch.randn(N, H, generator=g)
    masks = (torch.rand(T * N, generator=g) >= p_done).float()
    dout = torch.randn(T * N, H, generator=g)

    # fp64 reference with autograd
    gi_r = gi.double().requires_grad_(True)
    W_r = Whh.double().requires_grad_(True)
    b_r = bhh.double()
    h = h0.double()
    outs, hms = [], []
    for t in range(T):
        hm = h * masks[t * N:(t + 1) * N, None].double()
        hms.append(hm)
        gh = hm @ W_r.t() + b_r
        gt = gi_r[t * N:(t + 1) * N]
        r = torch.sigmoid(gt[:, :H] + gh[:, :H])
        z = torch.sigmoid(gt[:, H:2 * H] + gh[:, H:2 * H])
        n = torch.tanh(gt[:, 2 * H:] + r * gh[:, 2 * H:])
        h = (1 - z) * n + z * hm
        outs.append(h)
    out_ref = torch.cat(outs)
    (out_ref * dout.double()).sum().backward()

    dev = DEV
    dz = lambda *s: torch.zeros(*s, device=dev)
    bufs = dict(gi=gi.to(dev), gh=dz(T * N, 3 * H), hm=dz(T * N, H), h0=h0.to(dev), masks=masks.to(dev),
                out=dz(T * N, H), r=dz(T * N, H), z=dz(T * N, H), n=dz(T * N, H), dout=dout.to(dev),
                dgi=dz(T * N, 3 * H), dgh=dz(T * N, 3 * H), dcarry=dz(N, H))
    W_d, b_d = Whh.to(dev), bhh.to(dev)
    d = E.GruSeqDesc()
    for k, v in bufs.items():
        setattr(d, k, v.data_ptr())
    d.T, d.N, d.H, d.t = T, N, H, 0
    d.W_hh, d.b_hh = W_d.data_ptr(), b_d.data_ptr()
    lib = _native.lib()
    st = torch.cuda.current_stream().cuda_stream
    _native.check(lib.a2cb_gru_seq(0, ctypes.addressof(d), st), "gru_seq fwd")
    _native.check(lib.a2cb_gru_seq(1, ctypes.addressof(d), st), "gru_seq bwd")
    torch.cuda.synchronize()

    def rel(a, b):
        b = b.float()
        return float((a.cpu() - b).norm() / (b.norm() + 1e-30))
    assert rel(bufs["out"], out_ref.detach()) <= 2e-6
    assert rel(bufs["hm"], torch.cat(hms).detach()) <= 2e-6
    assert rel(bufs["dgi"], gi_r.grad) <= 5e-6
    # dW_hh = sum_t dgh_t^T hm_t  (the weight-gradient GEMM consumes dgh and hm)
    dW = bufs["dgh"].double().t() @ bufs["hm"].double()
    assert rel(dW.float(), W_r.grad) <= 5e-6


def test_backward_products_with_the_engines_own_relu_masks_match_fp64():
    """Pins the argument behind the gradient allowance of the tensor-core engine (tests/test_update_gpu.py): its forward
    pre-activations differ from an exact-fp32 forward by ~2e-6 relative, which can flip the relu' bit of the few
    activations that close to zero; everything ELSE in the backward pass -- the 3xTF32 data / weight gradient products
    and their split folds -- must agree with an fp64 reference to 1e-5 once both use the SAME masks.  The fp64 reference
    below takes its relu masks from the engine's own activations (a > 0) instead of its own pre-activations."""
    from a2c_ppo_acktr.model import Policy

    class Discrete(object):
        n = 6
    _native.set_gemm_mode(2)
    torch.manual_seed(1)
    pol = Policy((4, 84, 84), Discrete()).to(DEV)
    B = 256                                                    # a C3 minibatch
    g = torch.Generator().manual_seed(3)
    obs_u8 = torch.randint(0, 256, (B, 4, 84, 84), generator=g, dtype=torch.uint8)
    action = torch.randint(0, 6, (B, 1), generator=g)
    cv, cl, ce = torch.randn(B, 1, generator=g) / B, torch.randn(B, 1, generator=g) / B, 0.01
    v2, lp2, ent2, _ = pol.evaluate_actions(obs_u8.to(DEV), torch.zeros(B, 1, device=DEV), torch.ones(B, 1, device=DEV),
                                            action.to(DEV))
    rt = pol.runtime()
    assert ("b%d_a1" % B) in rt.arena.bufs, "this test reads the TMA-fed engine's NHWC activations"
    def mask(name, shape):                                     # NHWC x plane -> NCHW {0, 1} fp64
        a = rt.arena.bufs["b%d_%s" % (B, name)][:B * shape[0] * shape[1] * shape[2]].view(B, *shape)
        return (a > 0).permute(0, 3, 1, 2).double().cpu()
    m1, m2, m3 = mask("a1", (20, 20, 32)), mask("a2", (9, 9, 64)), mask("a3", (7, 7, 32))
    mh = (rt.arena.bufs["b%d_h" % B][:B * 512].view(B, 512) > 0).double().cpu()
    for p in pol.parameters():
        p.grad.zero_()
    ((v2 * cv.to(DEV)).sum() + (lp2 * cl.to(DEV)).sum() + ce * ent2).backward()
    ref = {k: v.detach().cpu().double().requires_grad_(True) for k, v in pol.state_dict().items()}
    x = obs_u8.double() / 255.0
    a = F.conv2d(x, ref["base.main.0.weight"], ref["base.main.0.bias"], stride=4) * m1
    a = F.conv2d(a, ref["base.main.2.weight"], ref["base.main.2.bias"], stride=2) * m2
    a = F.conv2d(a, ref["base.main.4.weight"], ref["base.main.4.bias"], stride=1) * m3
    h = F.linear(a.reshape(B, -1), ref["base.main.7.weight"], ref["base.main.7.bias"]) * mh
    value = F.linear(h, ref["base.critic_linear.weight"], ref["base.critic_linear.bias"])
    dist = torch.distributions.Categorical(logits=F.linear(h, ref["dist.linear.weight"], ref["dist.linear.bias"]))
    logp = dist.log_prob(action.squeeze(-1)).unsqueeze(-1)
    ((value * cv.double()).sum() + (logp * cl.double()).sum() + ce * dist.entropy().mean()).backward()
    for k, p in pol.named_parameters():
        want = ref[k].grad
        err = float((p.grad.cpu().double() - want).norm() / want.norm().clamp_min(1e-30))
        assert err <= 1e-5, (k, err)
