"""Builds liba2cb.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

Usage:  python build.py [--force] [--verbose]
The library lands next to this file so that `gpurun` ships it to the GPU box.
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "liba2cb.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
          "-Xptxas", "-warn-spills"]


def _nccl_paths():
    """torch-bundled NCCL (header + libnccl.so.2); the library is resolved at load time."""
    try:
        import nvidia.nccl as n
        base = os.path.dirname(n.__file__) if getattr(n, "__file__", None) else list(n.__path__)[0]
        return os.path.join(base, "include"), os.path.join(base, "lib")
    except Exception:
        return None, None


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def headers_mtime():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(HERE, "..", "include", "a2cb.h"))
    return max(os.path.getmtime(h) for h in hs)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    hm = headers_mtime()
    inc, libdir = _nccl_paths()
    extra_inc = ["-I", inc] if inc else []
    jobs = []
    objs = []
    for s in sources():
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s[:-3] + ".o")
        objs.append(obj)
        if (not force and os.path.exists(obj) and os.path.getmtime(obj) >= os.path.getmtime(src)
                and os.path.getmtime(obj) >= hm):
            continue
        cmd = [NVCC] + ARCH + CFLAGS + extra_inc + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, r

    if jobs:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for cmd, r in ex.map(run, jobs):
                out = (r.stdout + r.stderr).strip()
                if r.returncode != 0:
                    sys.stderr.write(" ".join(cmd) + "\n" + out + "\n")
                    raise RuntimeError("nvcc failed for " + cmd[-3])
                if out and (verbose or "warning" in out.lower() or "spill" in out.lower()):
                    print(out)
    if jobs or force or not os.path.exists(LIB):
        # NCCL is resolved with dlopen at first use (csrc/nccl_glue.cu): no link-time dependency
        link = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart", "-ldl"]
        r = subprocess.run(link, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
