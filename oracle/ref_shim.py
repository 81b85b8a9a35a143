"""Import shim for the UNMODIFIED reference (authoring container only).

The reference's modules import gym / stable_baselines3 / h5py at module scope
(a2c_ppo_acktr/utils.py:7 -> envs.py:3-17) although nothing on the update path
uses them, and algo/kfac.py:208-211 calls `torch.symeig`, which modern torch
removed.  This shim registers empty stand-in modules and maps symeig onto
`torch.linalg.eigh(UPLO='U')` (symeig's default triangle).  Nothing of the
reference is copied; it is imported from where it lies.

Used by tests/golden/make_golden.py and by the optional live cross-checks in
tests/ (skipped when /root/reference is absent, e.g. on the GPU box).
"""
import os
import sys
import types

REFERENCE_DIR = os.environ.get("A2CB_REFERENCE_DIR", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_DIR, "a2c_ppo_acktr"))


class _Dummy(object):
    def __init__(self, *a, **k):
        pass


class RunningMeanStd(object):
    """Functional stand-in for stable_baselines3.common.running_mean_std.RunningMeanStd (an ABSENT third-party
    dependency of reference algo/gail.py:9,30,109; the reference pins no version, requirements.txt:1-5).  Restates the
    published algorithm of that class (sb3 1.x / 2.x, unchanged since OpenAI baselines): parallel-variance merge of
    batch moments (Chan et al.), all float64, count starts at epsilon = 1e-4, mean 0, var 1."""

    def __init__(self, epsilon=1e-4, shape=()):
        import numpy as np
        self.mean = np.zeros(shape, np.float64)
        self.var = np.ones(shape, np.float64)
        self.count = epsilon

    def update(self, arr):
        import numpy as np
        self.update_from_moments(np.mean(arr, axis=0), np.var(arr, axis=0), arr.shape[0])

    def update_from_moments(self, batch_mean, batch_var, batch_count):
        import numpy as np
        delta = batch_mean - self.mean
        tot = self.count + batch_count
        new_mean = self.mean + delta * batch_count / tot
        m2 = self.var * self.count + batch_var * batch_count + np.square(delta) * self.count * batch_count / tot
        self.mean, self.var, self.count = new_mean, m2 / tot, tot


def _stub(name, attrs=()):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    for a in attrs:
        if not hasattr(m, a):
            setattr(m, a, type(a, (_Dummy,), {}))
    return m


def install():
    """Makes `import a2c_ppo_acktr` resolve to the reference; returns its modules."""
    if not available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_DIR)
    import torch

    gym = _stub("gym", ["Wrapper", "ObservationWrapper", "Env"])
    spaces = _stub("gym.spaces", ["Box", "Discrete", "MultiBinary"])
    gym.spaces = spaces
    _stub("gym.spaces.box", ["Box"])
    _stub("gym.wrappers")
    _stub("gym.wrappers.clip_action", ["ClipAction"])
    _stub("stable_baselines3")
    _stub("stable_baselines3.common")
    _stub("stable_baselines3.common.atari_wrappers",
          ["ClipRewardEnv", "EpisodicLifeEnv", "FireResetEnv", "MaxAndSkipEnv", "NoopResetEnv", "WarpFrame"])
    _stub("stable_baselines3.common.monitor", ["Monitor"])
    _stub("stable_baselines3.common.vec_env", ["DummyVecEnv", "SubprocVecEnv", "VecEnvWrapper"])
    _stub("stable_baselines3.common.vec_env.vec_normalize", ["VecNormalize"])
    _stub("stable_baselines3.common.running_mean_std").RunningMeanStd = RunningMeanStd
    _stub("h5py")

    if not hasattr(torch, "symeig") or getattr(torch.symeig, "_a2cb_shim", False) is False:
        def symeig(a, eigenvectors=False, upper=True):
            d, q = torch.linalg.eigh(a, UPLO='U' if upper else 'L')
            return d, q
        symeig._a2cb_shim = True
        torch.symeig = symeig

    # drop any previously imported product package of the same name
    for k in [k for k in sys.modules if k == "a2c_ppo_acktr" or k.startswith("a2c_ppo_acktr.")]:
        mod = sys.modules[k]
        f = getattr(mod, "__file__", "") or ""
        if not f.startswith(REFERENCE_DIR):
            del sys.modules[k]
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    import importlib
    mods = {}
    for name in ("storage", "model", "distributions", "utils", "algo", "algo.ppo", "algo.a2c_acktr", "algo.kfac",
                 "algo.gail"):
        mods[name] = importlib.import_module("a2c_ppo_acktr." + name)
    return mods


def uninstall():
    """Removes the reference from sys.modules / sys.path again."""
    for k in [k for k in sys.modules if k == "a2c_ppo_acktr" or k.startswith("a2c_ppo_acktr.")]:
        del sys.modules[k]
    while REFERENCE_DIR in sys.path:
        sys.path.remove(REFERENCE_DIR)


class Discrete(object):
    """Action-space stand-in; the reference dispatches on the class NAME only."""
    def __init__(self, n):
        self.n = n


class Box(object):
    def __init__(self, dim):
        self.shape = (dim,)
