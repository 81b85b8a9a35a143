"""CPU: host logic of the actor step and the ingest path -- descriptor pointer scanning / patching (`ActPlan`),
content-addressed device tables (`Arena.const`), and the argument checks of `RolloutStorage.stage_host`."""
import ctypes

import numpy as np
import pytest
import torch

from a2c_ppo_acktr import _engine, _native, _tcx
from a2c_ppo_acktr.storage import RolloutStorage


class _Inner(ctypes.Structure):
    _fields_ = [("x", ctypes.c_void_p), ("lo", ctypes.c_void_p), ("n", ctypes.c_int64)]


class _Outer(ctypes.Structure):
    _fields_ = [("a", _Inner), ("arr", _Inner * 2), ("p", ctypes.c_void_p), ("count", ctypes.c_int32), ("pad", ctypes.c_int32),
                ("vals", ctypes.c_int64 * 3)]


def test_pointer_sites_finds_nested_fields_and_patching_moves_them():
    d = _Outer()
    base, size = 0x7000_0000, 4096
    d.a.x, d.a.lo = base + 16, 0x1234                       # inside / outside
    d.arr[0].x, d.arr[1].lo = base + 4095, base + size        # last byte inside / one past the end
    d.p = base
    d.vals[0] = base + 8                                      # an integer that merely looks like a pointer: not scanned
    sites = []
    _engine._pointer_sites(d, base, base + size, sites)
    got = sorted((f, off) for _, f, off in sites)
    assert got == [("p", 0), ("x", 16), ("x", 4095)]
    plan = _engine.ActPlan.__new__(_engine.ActPlan)
    plan.sites, plan.home, plan.current, plan.graphs = {"obs": sites}, {"obs": base}, {"obs": base}, {}
    plan.point("obs", 0x9000_0000)
    assert (d.p, d.a.x, d.arr[0].x) == (0x9000_0000, 0x9000_0010, 0x9000_0FFF)
    assert d.a.lo == 0x1234 and d.arr[1].lo == base + size and d.vals[0] == base + 8
    plan.point("obs", base)                                   # and back home
    assert (d.p, d.a.x) == (base, base + 16)


def test_pointer_sites_walks_the_real_descriptors():
    """Every pointer field of the engine's descriptors is a c_void_p reachable by the scan (a pointer stored as a
    plain integer would silently escape patching)."""
    seen = []
    for cls in (_engine.LossDesc, _engine.GruDesc, _engine.GruSeqDesc, _engine.MlpDesc, _engine.SampleDesc, _tcx.FwdDesc,
                _tcx.WgradDesc, _tcx.S2dDesc, _tcx.NarrowFwdDesc, _engine.P.GemmDesc):
        d = cls()
        ptr_fields = [n for n, t in cls._fields_ if t is ctypes.c_void_p]
        for i, n in enumerate(ptr_fields):
            setattr(d, n, 0x5000_0000 + 8 * i)
        out = []
        _engine._pointer_sites(d, 0x5000_0000, 0x5000_0000 + 8 * len(ptr_fields), out)
        assert sorted(f for _, f, _ in out) == sorted(ptr_fields), cls.__name__
        seen.append(len(ptr_fields))
    v = _tcx.FwdDesc()
    v.A.x = 0x5000_0000
    out = []
    _engine._pointer_sites(v, 0x5000_0000, 0x5000_0001, out)
    assert [(f, o) for _, f, o in out] == [("x", 0)] and sum(seen) > 60


def test_arena_const_is_content_addressed_and_never_replaced():
    a = _engine.Arena(torch.device("cpu"))
    t1 = a.const(np.arange(5, dtype=np.int32))
    t2 = a.const(np.arange(5, dtype=np.int32))
    t3 = a.const(np.arange(5, dtype=np.int64))
    t4 = a.const(np.arange(6, dtype=np.int32))
    assert t1.data_ptr() == t2.data_ptr() and len({t1.data_ptr(), t3.data_ptr(), t4.data_ptr()}) == 3
    assert t1.tolist() == [0, 1, 2, 3, 4]


def test_stage_host_argument_checks_and_cpu_storage_is_a_noop_for_waits():
    class Discrete(object):
        n = 6
    st = RolloutStorage(4, 3, (4, 84, 84), Discrete(), 1, obs_dtype=torch.uint8)
    st.wait_staged()                                          # nothing staged: no device touched
    st.after_update()
    with pytest.raises(_native.NativeError):
        st.stage_host({"rewards": torch.zeros(4, 3, 1)})      # buffer lives on the CPU: no fallback
    assert st._staged_obs is None
