"""GPU parity of the returns scan and the minibatch gather (through the C ABI) against the
oracle and the reference-produced golden files."""
import numpy as np
import pytest
import torch

import helpers
from oracle import minibatch as o_mb
from oracle import returns as o_ret

pytestmark = pytest.mark.gpu

from a2c_ppo_acktr import _native  # noqa: E402
from a2c_ppo_acktr.storage import RolloutStorage  # noqa: E402


class Discrete(object):
    def __init__(self, n):
        self.n = n


class Box(object):
    def __init__(self, d):
        self.shape = (d,)


DEV = "cuda:0"


def _scan(g, si, use_gae, ptl, schedule):
    r = torch.from_numpy(g["in%d_rewards" % si]).to(DEV)
    v = torch.from_numpy(g["in%d_values" % si]).to(DEV)
    m = torch.from_numpy(g["in%d_masks" % si]).to(DEV)
    b = torch.from_numpy(g["in%d_bad" % si]).to(DEV)
    nv = torch.from_numpy(g["in%d_next" % si]).to(DEV).reshape(-1).contiguous()
    out = torch.zeros_like(v)
    _native.returns_scan(r, v, out, m, b, nv, 0.99, 0.95, use_gae, ptl, schedule)
    torch.cuda.synchronize()
    return out.cpu().numpy(), v.cpu().numpy()


@pytest.mark.parametrize("si", range(7))
@pytest.mark.parametrize("use_gae", [0, 1])
@pytest.mark.parametrize("ptl", [0, 1])
def test_scan_env_schedule_bit_exact_vs_golden(si, use_gae, ptl):
    g = helpers.load("returns")
    out, v = _scan(g, si, use_gae, ptl, schedule=1)
    want = g["out%d_g%d_p%d_returns" % (si, use_gae, ptl)]
    T = want.shape[0] - 1
    upto = T if use_gae else T + 1
    assert np.array_equal(out[:upto], want[:upto])
    assert np.array_equal(v[-1], g["out%d_g%d_p%d_vlast" % (si, use_gae, ptl)])


@pytest.mark.parametrize("si", range(7))
@pytest.mark.parametrize("use_gae", [0, 1])
@pytest.mark.parametrize("ptl", [0, 1])
def test_scan_time_schedule_within_1e5(si, use_gae, ptl):
    g = helpers.load("returns")
    out, v = _scan(g, si, use_gae, ptl, schedule=2)
    want = g["out%d_g%d_p%d_returns" % (si, use_gae, ptl)]
    T = want.shape[0] - 1
    upto = T if use_gae else T + 1
    # north_star: GAE returns within 1e-5 relative (|d| <= 1e-5 * max(1, |ref|))
    err = np.abs(out[:upto] - want[:upto]) / np.maximum(1.0, np.abs(want[:upto]))
    assert err.max() <= 1e-5, err.max()
    assert np.array_equal(v[-1], g["out%d_g%d_p%d_vlast" % (si, use_gae, ptl)])


@pytest.mark.parametrize("T,N", [(128, 2048), (16, 4096), (3, 100003), (128, 8), (2048, 1), (0, 4)])
@pytest.mark.parametrize("use_gae,ptl", [(1, 1), (1, 0), (0, 1), (0, 0)])
def test_scan_vs_oracle_random(T, N, use_gae, ptl):
    gen = torch.Generator().manual_seed(T * 7 + N)
    r = torch.randn(T, N, 1, generator=gen)
    v = torch.randn(T + 1, N, 1, generator=gen)
    m = (torch.rand(T + 1, N, 1, generator=gen) > 0.02).float()
    b = (torch.rand(T + 1, N, 1, generator=gen) > 0.01).float()
    nv = torch.randn(N, 1, generator=gen)
    want, wv = o_ret.compute_returns(r.numpy(), v.numpy(), m.numpy(), b.numpy(), nv.numpy(), bool(use_gae), 0.99,
                                     0.95, bool(ptl))
    upto = T if use_gae else T + 1
    for schedule in (1, 2, 0):
        rd, vd, md, bd = r.to(DEV), v.to(DEV), m.to(DEV), b.to(DEV)
        out = torch.zeros(T + 1, N, 1, device=DEV)
        _native.returns_scan(rd, vd, out, md, bd, nv.to(DEV).reshape(-1).contiguous(), 0.99, 0.95, use_gae, ptl,
                             schedule)
        got = out.cpu().numpy()
        if schedule == 1:
            assert np.array_equal(got[:upto], want[:upto])
        else:
            err = np.abs(got[:upto] - want[:upto]) / np.maximum(1.0, np.abs(want[:upto]))
            assert err.size == 0 or err.max() <= 1e-5
        assert np.array_equal(vd.cpu().numpy()[-1], wv[-1])


def test_scan_linearity_property_full_size():
    # size-independent property at the C5 shape: non-GAE returns are linear in (rewards, next_value)
    T, N = 128, 2048
    gen = torch.Generator().manual_seed(1)
    m = (torch.rand(T + 1, N, 1, generator=gen) > 0.01).float().to(DEV)
    b = torch.ones(T + 1, N, 1, device=DEV)
    v = torch.zeros(T + 1, N, 1, device=DEV)
    outs = []
    rs = [torch.randn(T, N, 1, generator=gen).double() for _ in range(2)]
    nvs = [torch.randn(N, generator=gen).double() for _ in range(2)]
    rs.append(rs[0] + rs[1])
    nvs.append(nvs[0] + nvs[1])
    for r, nv in zip(rs, nvs):
        out = torch.zeros(T + 1, N, 1, device=DEV)
        _native.returns_scan(r.float().to(DEV), v, out, m, b, nv.float().to(DEV), 0.99, 0.95, 0, 0, 1)
        outs.append(out.double())
    err = (outs[0] + outs[1] - outs[2]).abs().max().item()
    assert err < 1e-4


def _mk_storage(T, N, obs_shape, space, H, obs_dtype, gen):
    st = RolloutStorage(T, N, obs_shape, space, H, obs_dtype=obs_dtype)
    if obs_dtype == torch.uint8:
        st.obs.copy_(torch.randint(0, 256, st.obs.shape, generator=gen, dtype=torch.uint8))
    else:
        st.obs.copy_(torch.randn(st.obs.shape, generator=gen))
    for name in ("recurrent_hidden_states", "rewards", "value_preds", "returns", "action_log_probs", "masks"):
        t = getattr(st, name)
        t.copy_(torch.randn(t.shape, generator=gen))
    if st.actions.dtype == torch.int64:
        st.actions.copy_(torch.randint(0, 6, st.actions.shape, generator=gen))
    else:
        st.actions.copy_(torch.randn(st.actions.shape, generator=gen))
    return st


def _as_dict(st):
    return {n: getattr(st, n).clone() for n in ("obs", "recurrent_hidden_states", "rewards", "value_preds", "returns",
                                                "action_log_probs", "actions", "masks")}


@pytest.mark.parametrize("tag,kw", [("ffA", dict(num_mini_batch=4)), ("ffB", dict(mini_batch_size=5))])
def test_feed_forward_generator_vs_golden(tag, kw):
    g = helpers.load("generators")
    sp = Discrete(6) if tag == "ffA" else Box(2)
    T, N = g[tag + "_st_rewards"].shape[:2]
    st = RolloutStorage(T, N, (2, 3, 3), sp, 5)
    for n in ("obs", "recurrent_hidden_states", "rewards", "value_preds", "returns", "action_log_probs", "actions",
              "masks"):
        getattr(st, n).copy_(torch.from_numpy(g["%s_st_%s" % (tag, n)]))
    st.to(DEV)
    adv = torch.from_numpy(g[tag + "_adv"]).to(DEV)
    torch.manual_seed(7)
    batches = list(st.feed_forward_generator(adv, **kw))
    assert len(batches) == int(g[tag + "_nb"])
    names = ["obs", "hxs", "actions", "values", "returns", "masks", "oldlp", "adv"]
    for bi, tup in enumerate(batches):
        for name, x in zip(names, tup):
            want = g["%s_b%d_%s" % (tag, bi, name)]
            assert x.dtype == torch.from_numpy(want).dtype
            assert np.array_equal(x.cpu().numpy(), want), (tag, bi, name)
    torch.manual_seed(7)
    assert next(iter(st.feed_forward_generator(None, **kw)))[-1] is None


def test_recurrent_generator_vs_golden():
    g = helpers.load("generators")
    T, N = g["rec_st_rewards"].shape[:2]
    st = RolloutStorage(T, N, (2, 3, 3), Discrete(6), 5)
    for n in ("obs", "recurrent_hidden_states", "rewards", "value_preds", "returns", "action_log_probs", "actions",
              "masks"):
        getattr(st, n).copy_(torch.from_numpy(g["rec_st_%s" % n]))
    st.to(DEV)
    adv = torch.from_numpy(g["rec_adv"]).to(DEV)
    torch.manual_seed(7)
    batches = list(st.recurrent_generator(adv, 4))
    assert len(batches) == int(g["rec_nb"])
    names = ["obs", "hxs", "actions", "values", "returns", "masks", "oldlp", "adv"]
    for bi, tup in enumerate(batches):
        for name, x in zip(names, tup):
            assert np.array_equal(x.cpu().numpy(), g["rec_b%d_%s" % (bi, name)]), (bi, name)


@pytest.mark.parametrize("obs_dtype", [torch.uint8, torch.float32])
@pytest.mark.parametrize("T,N,nmb", [(128, 8, 4), (5, 16, 1), (16, 33, 7)])
def test_gather_atari_rows_bit_exact_vs_oracle(obs_dtype, T, N, nmb):
    gen = torch.Generator().manual_seed(T + N)
    st = _mk_storage(T, N, (4, 84, 84), Discrete(6), 1, obs_dtype, gen)
    ro = _as_dict(st)
    adv = torch.randn(T, N, 1, generator=gen)
    st.to(DEV)
    torch.manual_seed(11)
    got = list(st.feed_forward_generator(adv.to(DEV), nmb))
    torch.manual_seed(11)
    chunks = o_mb.feed_forward_index_chunks(T, N, nmb)
    assert len(got) == len(chunks)
    for tup, idx in zip(got, chunks):
        want = o_mb.gather_feed_forward(ro, idx, adv)
        for x, w in zip(tup, want):
            assert torch.equal(x.cpu(), w)


def test_recurrent_gather_wide_hidden_vs_oracle():
    T, N, H = 16, 12, 512
    gen = torch.Generator().manual_seed(2)
    st = _mk_storage(T, N, (4, 84, 84), Discrete(6), H, torch.uint8, gen)
    ro = _as_dict(st)
    adv = torch.randn(T, N, 1, generator=gen)
    st.to(DEV)
    torch.manual_seed(5)
    got = list(st.recurrent_generator(adv.to(DEV), 3))
    torch.manual_seed(5)
    chunks = o_mb.recurrent_env_chunks(N, 3)
    for tup, env in zip(got, chunks):
        want = o_mb.gather_recurrent(ro, env, adv)
        for x, w in zip(tup, want):
            assert torch.equal(x.cpu(), w)


def test_gather_permutation_roundtrip_full_size():
    # size-independent property at the C3 shape: gathering with a permutation and then with its
    # inverse restores the rows; a checksum of the gathered rows equals the checksum of the source
    T, N = 128, 8
    gen = torch.Generator().manual_seed(3)
    src = torch.randint(0, 256, (T * N, 4 * 84 * 84), generator=gen, dtype=torch.uint8).to(DEV)
    perm = torch.randperm(T * N, generator=gen)
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(T * N)
    a = torch.empty_like(src)
    b = torch.empty_like(src)
    _native.gather_rows([(src, a)], perm.to(DEV), T * N)
    _native.gather_rows([(a, b)], inv.to(DEV), T * N)
    assert torch.equal(b, src)
    assert int(a.long().sum()) == int(src.long().sum())


def test_gather_empty_and_errors():
    src = torch.zeros(4, 8, device=DEV)
    dst = torch.zeros(0, 8, device=DEV)
    _native.gather_rows([(src, dst)], torch.zeros(0, dtype=torch.int64, device=DEV), 0)
    with pytest.raises(_native.NativeError):
        _native.gather_rows([(src, torch.zeros(2, 8, device=DEV))], torch.zeros(2, dtype=torch.int64), 2)
