"""CPU-only: the op lists `_tcx.CnnNet` builds for the TMA-fed engine (tap tables, parity views, box / tile
geometry, weight-image and row-offset tables, the once-per-update frame re-layout with image gather),
evaluated with the numpy emulator `_tcx.emulate`, reproduce torch's CNNBase forward and autograd gradients."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from a2c_ppo_acktr import _tcx


class _Spec(object):
    hidden = 512


class _Layout(object):
    def __init__(self, entries):
        self.entries = entries


class _FakeBuilder(object):
    """The slice of _engine.NetBuilder that CnnNet touches."""

    def __init__(self, B, entries, obs_code, gather, full_rows):
        self.B, self.spec, self.L = B, _Spec(), _Layout(entries)
        self.obs_code, self.gather, self.full_rows, self.obs = obs_code, gather, full_rows, "obs"
        self.sizes, self.t = {}, "t_"

    def n(self, name):
        return self.t + name

    def buf(self, name, numel):
        self.sizes[self.n(name)] = int(numel)
        return (self.n(name), 0)


@pytest.mark.parametrize("gather_full", [False, True, "chunk"])
@pytest.mark.parametrize("u8", [True, False])
def test_cnn_trunk_ops_match_torch(u8, gather_full):
    torch.manual_seed(0)
    Bn, rows = 3, 7
    obs_u8 = torch.randint(0, 256, (rows, 4, 84, 84), dtype=torch.uint8)
    idx = np.array([5, 0, 3])
    shapes = [("base.main.0", (32, 4, 8, 8)), ("base.main.2", (64, 32, 4, 4)), ("base.main.4", (32, 64, 3, 3)),
              ("base.main.7", (512, 1568))]
    ws = [torch.randn(*s) * 0.05 for _, s in shapes]
    bs = [torch.randn(s[0]) * 0.1 for _, s in shapes]
    for t in ws + bs:
        t.requires_grad_(True)
    x = obs_u8[idx].float() / 255.0
    a1 = F.relu(F.conv2d(x, ws[0], bs[0], stride=4))
    a2 = F.relu(F.conv2d(a1, ws[1], bs[1], stride=2))
    a3 = F.relu(F.conv2d(a2, ws[2], bs[2], stride=1))
    h = F.relu(F.linear(a3.reshape(Bn, -1), ws[3], bs[3]))
    Rm = torch.randn_like(h)
    (h * Rm).sum().backward()

    entries, o = {}, 0
    for (name, s), w, b in zip(shapes, ws, bs):
        entries[name + ".weight"] = (o, s)
        entries[name + ".bias"] = (o + w.numel(), (s[0],))
        o += w.numel() + b.numel()
    Pbuf = np.concatenate([np.concatenate([w.detach().numpy().ravel(), b.detach().numpy().ravel()])
                           for w, b in zip(ws, bs)]).astype(np.float32)
    nb = _FakeBuilder(Bn, entries, 1 if u8 else 2, True, rows if gather_full else 0)
    net = _tcx.CnnNet(nb)
    fwd, href = net.forward()
    gh = nb.buf("gh", Bn * 512)
    gh_lo = nb.buf("gh_lo", Bn * 512)
    bwd = net.backward(gh, gh_lo)
    bufs = {"obs": (obs_u8.numpy().ravel() if u8 else obs_u8.float().numpy().ravel()), "P": Pbuf, "G": np.zeros_like(Pbuf)}
    for name, numel in nb.sizes.items():
        bufs[name] = np.zeros(numel, np.float32)
    assert bool(net.prologue) == bool(gather_full) and bool(net.prologue_chunk) == bool(gather_full)
    if gather_full == "chunk":
        # per-minibatch re-layout: only the minibatch's rows are converted, each into its own slot of the full tensor
        for op in net.prologue_chunk:
            _tcx.emulate(op, bufs, idx)
        x0 = bufs["x0_full"].reshape(rows, -1)
        assert all((np.abs(x0[r]).sum() > 0) == (r in idx) for r in range(rows))
    else:
        for op in net.prologue:
            _tcx.emulate(op, bufs)
    for op in fwd:
        _tcx.emulate(op, bufs, idx)
    nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous().detach().numpy()
    tol = dict(rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(bufs["t_a1"].reshape(Bn, 20, 20, 32), nhwc(a1), **tol)
    np.testing.assert_allclose(bufs["t_a2"].reshape(Bn, 9, 9, 64), nhwc(a2), **tol)
    np.testing.assert_allclose(bufs["t_a3"].reshape(Bn, 7, 7, 32), nhwc(a3), **tol)
    np.testing.assert_allclose(bufs[href[0]].reshape(Bn, 512), h.detach().numpy(), **tol)

    bufs["t_gh"][:] = (Rm * (h > 0)).detach().numpy().ravel()           # gradient w.r.t. h, relu' applied
    for op in bwd:
        _tcx.emulate(op, bufs, idx)
    G = bufs["G"]
    for (name, s), w, b in zip(shapes, ws, bs):
        ow, _ = entries[name + ".weight"]
        ob, _ = entries[name + ".bias"]
        np.testing.assert_allclose(G[ow:ow + w.numel()].reshape(s), w.grad.numpy(), rtol=2e-4, atol=2e-5, err_msg=name + ".weight")
        np.testing.assert_allclose(G[ob:ob + b.numel()], b.grad.numpy(), rtol=2e-4, atol=2e-5, err_msg=name + ".bias")


def test_hoist_preps_merges_weight_images():
    nb = _FakeBuilder(2, {k: (0, ()) for k in ("base.main.0.weight", "base.main.0.bias", "base.main.2.weight", "base.main.2.bias",
                                               "base.main.4.weight", "base.main.4.bias", "base.main.7.weight", "base.main.7.bias")},
                      1, False, 0)
    net = _tcx.CnnNet(nb)
    fwd, _ = net.forward()
    bwd = net.backward(nb.buf("gh", 1024), nb.buf("gh_lo", 1024))
    n_prep = sum(1 for o in fwd + bwd if isinstance(o, _tcx.TcxOp) and o.kind == _tcx.OP_TCX_PREP)
    f2, b2 = _tcx.hoist_preps([fwd, bwd], "t_")
    assert n_prep == 8          # 4 forward (+ the bf16 image of conv1 for uint8 frames) + 3 data-gradient images
    assert f2[0].kind == _tcx.OP_TCX_PREP_MULTI
    assert not any(isinstance(o, _tcx.TcxOp) and o.kind == _tcx.OP_TCX_PREP for o in f2 + b2)
    assert len(f2) + len(b2) == len(fwd) + len(bwd) - n_prep + 1
