"""Benchmark of the rollout-batch policy-update hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C3]

One "step" = one pass of the hot path over one synthetic rollout:
    RolloutStorage.compute_returns -> PPO.update -> after_update        (reference main.py:157-162)
`value` is update transitions/s with the rollout already resident in HBM; `e2e` is the
same metric through the public a2c_ppo_acktr API with HOST buffers: every step copies the
rollout from pinned host memory to the device storage and reads the three losses back.
Prints ONE JSON line (rank 0).  `--impl reference` times the CPU oracle port of the
reference (oracle/) on the host cores with the same config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "pytorch-a2c-ppo-acktr-gail_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import synthetic  # noqa: E402


class Discrete(object):
    def __init__(self, n):
        self.n = n


class Box(object):
    def __init__(self, d):
        self.shape = (d,)


def space(action):
    kind, n = action
    return Discrete(n) if kind == "discrete" else Box(n)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tensor_burst=p["bf16_tflops"], tensor_sustained=p["bf16_tflops_sustained"],
                    source="MEASURED_PEAKS.json")
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor_sustained=1400.0, source="fallback (B200_PROFILING.md)")


def measure_tf32_peak(dev, seconds=1.5):
    """Dense TF32 peak of THIS box, measured the way MEASURED_PEAKS.json measures bf16: cuBLAS fp32 matmul with TF32
    tensor cores on 8192^3 (2 N^3 flops), best of 10 (burst) and back to back for `seconds` (sustained)."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a = torch.randn(n, n, device=dev)
        b = torch.randn(n, n, device=dev)
        c = torch.empty(n, n, device=dev)
        for _ in range(3):
            torch.matmul(a, b, out=c)
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b, out=c)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(10, int(seconds * 1e3 / best))
        e0.record()
        for _ in range(reps):
            torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        flops = 2.0 * n ** 3
        return {"tf32_tflops": flops / best / 1e9, "tf32_tflops_sustained": flops * reps / e0.elapsed_time(e1) / 1e9,
                "how": "torch.matmul fp32 with allow_tf32 (cuBLAS TF32 tensor cores), 8192^3, best of 10 / %d back to back" % reps}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


# ---------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ---------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                c, m = float(parts[0]), float(parts[1])
            except ValueError:
                continue
            if t0 - 0.05 <= ts <= t1 + 0.2:
                sm.append(c)
                smax.append(m)
                for n, v in zip(names, parts[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        if not sm:
            for ts, line in self.lines[-3:]:
                parts = [p.strip() for p in line.split(",")]
                try:
                    sm.append(float(parts[0]))
                    smax.append(float(parts[1]))
                except Exception:
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------
# workload construction
# ---------------------------------------------------------------------------
ENVS_OVERRIDE = 0


def workload_config(name):
    cfg = dict(synthetic.CONFIGS[name])
    cfg["name"] = name
    if name == "C5":
        cfg["N"] = 256          # one GPU's shard of the 2048 environments (256 envs / GPU)
    if ENVS_OVERRIDE:
        cfg["N"] = ENVS_OVERRIDE
    return cfg


def run_ours(args):
    import torch.distributed as dist
    from a2c_ppo_acktr import _native, algo
    from a2c_ppo_acktr.model import Policy
    from a2c_ppo_acktr.storage import RolloutStorage

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs a torchrun launch with %d ranks" % (args.gpus, args.gpus))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    from a2c_ppo_acktr import utils as a2c_utils
    # N > 1: every rank pins itself to its GPU's NUMA node before the pinned rollout is allocated.  (Not at N = 1: the
    # same process later times the CPU baseline on ALL host cores, and worker threads inherit the mask.)
    bound = None if (args.no_numa_bind or world == 1) else a2c_utils.bind_host_to_device_node(local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL's banner / logs: stdout carries ONE JSON line
        dist.init_process_group("nccl", device_id=dev)

    cfg = workload_config(args.workload)
    T, Nloc = cfg["T"], cfg["N"]                    # weak scaling: every rank owns cfg["N"] environments
    N_total = Nloc * world
    env_range = (rank * Nloc, (rank + 1) * Nloc)

    torch.manual_seed(1)
    policy = Policy(cfg["obs_shape"], space(cfg["action"]), base_kwargs={"recurrent": cfg["recurrent"]})
    policy.to(dev)
    H = policy.recurrent_hidden_state_size
    obs_dtype = torch.uint8 if cfg["kind"] == "atari" else torch.float32

    # ---- synthetic rollout: frames on the host, policy outputs from Policy.act on the device ----
    # Atari-shaped workloads use the frame-ring format by default (--stacked-frames: the round-1 layout): every 84 x 84
    # frame is generated and stored ONCE, observation t is the rolling stack of frames t .. t + 3 with the stack cleared
    # where an episode ended (reference envs.py:218-259) -- a quarter of the bytes on the host link and in HBM.
    torch.manual_seed(2 + rank)
    # (PPO workloads: the update reads the ring in place.  A2C / ACKTR run ONE full-batch pass per update on a few hundred
    # rows and would have to materialise the stacks first, so they keep the stacked layout.)
    ring = cfg["kind"] == "atari" and cfg["algo"] == "ppo" and not args.stacked_frames
    if ring:
        C = cfg["obs_shape"][0]
        fr_np = synthetic.atari_frames(T + C, N_total, 0, env_range, (1,) + tuple(cfg["obs_shape"][1:]))
        fr_np = fr_np.reshape((T + C, Nloc) + tuple(cfg["obs_shape"][1:]))
        rew_np, mask_np, bad_np = synthetic.atari_signals(T, N_total, 0, env_range)
        valid_np = np.ones((T + 1, Nloc), np.uint8)
        valid_np[0] = C
        for t in range(T):
            valid_np[t + 1] = np.where(mask_np[t + 1, :, 0] == 0, 1, np.minimum(valid_np[t].astype(np.int64) + 1, C))
        host = {"frames": torch.from_numpy(fr_np).pin_memory(), "stack_valid": torch.from_numpy(valid_np).pin_memory()}
        obs_key = ("frames", "stack_valid")
    else:
        obs_np, rew_np, mask_np, bad_np = synthetic.rollout_inputs(cfg, 0, env_range, N_total, obs_uint8=(obs_dtype == torch.uint8))
        host = {"obs": torch.from_numpy(obs_np).pin_memory()}
        obs_key = ("obs",)
    host.update({
        "rewards": torch.from_numpy(rew_np).pin_memory(),
        "masks": torch.from_numpy(mask_np).pin_memory(),
        "bad_masks": torch.from_numpy(bad_np).pin_memory(),
    })
    st = RolloutStorage(T, Nloc, cfg["obs_shape"], space(cfg["action"]), H, obs_dtype=obs_dtype, frame_ring=ring)
    st.to(dev)
    for k in obs_key + ("rewards", "masks", "bad_masks"):
        getattr(st, k).copy_(host[k])
    with torch.no_grad():
        for t in range(T):
            v, a, lp, h = policy.act(st.obs[t], st.recurrent_hidden_states[t], st.masks[t])
            st.value_preds[t].copy_(v)
            st.actions[t].copy_(a)
            st.action_log_probs[t].copy_(lp)
            st.recurrent_hidden_states[t + 1].copy_(h)
        next_value = policy.get_value(st.obs[T], st.recurrent_hidden_states[T], st.masks[T]).clone()
    for k in ("actions", "action_log_probs", "value_preds", "recurrent_hidden_states"):
        host[k] = getattr(st, k).cpu().pin_memory()
    host["next_value"] = next_value.cpu().pin_memory()
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())
    rollout_keys = obs_key + ("rewards", "masks", "bad_masks", "actions", "action_log_probs", "value_preds",
                              "recurrent_hidden_states")

    if cfg["algo"] == "ppo":
        agent = algo.PPO(policy, cfg["clip_param"], cfg["ppo_epoch"], cfg["num_mini_batch"], cfg["value_loss_coef"],
                         cfg["entropy_coef"], lr=cfg["lr"], eps=cfg["eps"], max_grad_norm=cfg["max_grad_norm"])
    elif cfg["algo"] == "acktr":
        agent = algo.A2C_ACKTR(policy, cfg["value_loss_coef"], cfg["entropy_coef"], acktr=True)
    else:
        agent = algo.A2C_ACKTR(policy, cfg["value_loss_coef"], cfg["entropy_coef"], lr=cfg["lr"], eps=cfg["eps"],
                               alpha=cfg["alpha"], max_grad_norm=cfg["max_grad_norm"])
    if world > 1:
        agent.set_distributed(world, rank, env_range, N_total)
    if hasattr(agent, "recurrent_sharding") and args.recurrent_sharding:
        agent.recurrent_sharding = args.recurrent_sharding

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def upload():
        for k in rollout_keys:
            getattr(st, k).copy_(host[k], non_blocking=True)
        return host["next_value"].to(dev, non_blocking=True)

    def upload_e2e():
        """The public ingest call: narrow tensors now, frames pulled by update() on a side stream (recurrent PPO:
        per minibatch environment set, so that the upload overlaps the compute of earlier minibatches)."""
        st.stage_host({k: host[k] for k in rollout_keys})
        return host["next_value"].to(dev, non_blocking=True)

    def hot_path(nv):
        st.compute_returns(nv, cfg["use_gae"], cfg["gamma"], cfg["gae_lambda"], cfg["use_proper_time_limits"])
        out = agent.update(st)
        st.after_update()
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n_steps, e2e, profile_once=False):
        """Device-timed steps (CUDA events on the launching stream), L2 flushed between steps."""
        total_ms = 0.0
        losses = None
        for i in range(n_steps):
            torch.manual_seed(7 + i)
            if not e2e:
                nv = upload()                                  # restore the rollout (params keep training)
            flush.zero_()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            prof = profile_once and not e2e and i == 0 and os.environ.get("A2CB_PROFILE_STEP") == "1"
            if prof:
                torch.cuda.profiler.start()                    # ncu --profile-from-start off: one timed step
            e0.record()
            if e2e:
                nv = upload_e2e()
            losses = hot_path(nv)                              # update() ends with the D2H of the losses
            e1.record()
            if prof:
                torch.cuda.synchronize()
                torch.cuda.profiler.stop()
            barrier()
            ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            total_ms += ms.item()
            if rank == 0 and os.environ.get("A2CB_BENCH_VERBOSE") == "1":
                sys.stderr.write("step %d (%s): %.3f ms\n" % (i, "e2e" if e2e else "resident", ms.item()))
        return total_ms, losses

    timed(args.warmup, False)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    t0 = time.time()
    l0 = _native.launch_count()
    total_ms, losses = timed(args.steps, False, profile_once=True)
    launches = _native.launch_count() - l0
    t1 = time.time()
    clocks = sampler.stop(t0, t1)
    timed(1, True)
    e2e_ms, _ = timed(args.steps, True)
    # N > 1, recurrent PPO: the default forms the reference's minibatches (ONE global randperm per epoch) and migrates the
    # surplus environments of a minibatch between ranks.  The stratified opt-in scheme (balanced by construction, NOT the
    # reference's minibatch composition, no migration) is timed next to it, a few steps, as the no-migration bound.
    strat = None
    if world > 1 and cfg.get("recurrent") and getattr(agent, "recurrent_sharding", "") in ("global", "balanced"):
        default_mode = agent.recurrent_sharding
        agent.recurrent_sharding = "stratified"
        timed(2, False)
        k = max(3, min(args.steps, 5))
        s_ms, _ = timed(k, False)
        agent.recurrent_sharding = default_mode
        strat = {"value": T * Nloc * world / (s_ms / k / 1e3), "unit": "transitions/s", "ms_per_step": s_ms / k, "steps": k,
                 "note": "recurrent_sharding='stratified': balanced per-rank permutations, not the reference's minibatch "
                         "composition; shown to separate load imbalance from exchange cost"}

    # N > 1: after all those sharded updates every rank must hold bit-identical parameters (one rank sums each gradient
    # element, the clip + Adam step is redundant and identical) -- a cheap end-to-end check of the exchange at this N
    identical = None
    if world > 1:
        flat = policy._flat.detach().clone()
        ref = flat.clone()
        dist.broadcast(ref, 0)
        ok = torch.tensor([1 if (torch.equal(flat, ref) and bool(torch.isfinite(flat).all())) else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        identical = bool(ok.item())
        if not identical:
            raise SystemExit("bench: parameters diverged across ranks (or are not finite) after the sharded updates")

    transitions = T * Nloc * world
    ms_per_step = total_ms / args.steps
    value = transitions / (ms_per_step / 1e3)
    e2e_value = transitions / (e2e_ms / args.steps / 1e3)

    out = None
    if rank == 0:
        pk = peaks()
        # the TF32 peak is measured AFTER the per-kernel timings: seconds of dense GEMMs at the power cap pull the SM clock
        # down for a while, which would inflate every kernel time taken right after them
        roof, kernels = kernel_roofline(agent, st, cfg, dev, pk, lambda: measure_tf32_peak(dev)) if world == 1 else (None, None)
        scan = scan_roofline(dev, pk) if world == 1 else None
        out = {
            "metric": "%s update transitions/sec" % cfg["algo"].upper(),
            "value": value, "unit": "transitions/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s: %s" % (cfg["name"], describe(cfg)), "envs_per_gpu": Nloc, "num_steps": T,
                       "recurrent_sharding": (getattr(agent, "recurrent_sharding", None) if (world > 1 and cfg.get("recurrent")) else None),
                       "exchange": (None if world == 1 else
                                    "peer-memory all-reduce over NVLink mappings (csrc/p2p.cu)" if getattr(getattr(agent, "_comm", None), "p2p", None) is not None
                                    else "ncclAllReduce on a side stream" if getattr(agent, "_comm", None) is not None else "torch.distributed"),
                       "obs": ("uint8 frame ring (each 84x84 frame stored once, 4-stacks assembled on the device)" if ring
                               else "uint8 4x84x84 stacks" if obs_dtype == torch.uint8 else "fp32"), "l2": "flushed between steps (256 MiB write)",
                       "step": "compute_returns + update + after_update",
                       "host_cpus": ("%d cores of the GPU's NUMA node" % len(bound)) if bound else "unbound"},
            "e2e": {"value": e2e_value, "unit": "transitions/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": 12},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "losses": [float(x) for x in losses],
        }
        if strat:
            out["stratified_sharding"] = strat
        if identical is not None:
            out["ranks_bit_identical"] = identical
        if roof:
            out["roofline"] = roof
            out["kernels"] = kernels
        if scan:
            out["scan"] = scan
        if world == 1 and cfg["name"] == "C5" and Nloc == 256 and not args.no_parity:
            out["parity"] = golden_parity_check(cfg, None if ring else obs_np, dev)
        if world == 1 and not args.no_cpu_baseline:
            stacks_np = st.obs[:].cpu().numpy() if ring else obs_np       # the oracle port reads stacked fp32 observations
            out["cpu_baseline"] = cpu_baseline(cfg, (stacks_np, rew_np, mask_np, bad_np))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if out is not None:
        print(json.dumps(out))


def golden_parity_check(cfg, obs_np, dev):
    """First update of the benchmarked shape against the REFERENCE's losses: tests/golden/update_C5m.npz holds the
    policy outputs and losses the unmodified reference produced on this very rollout (same counter-based frames,
    128 x 256, one epoch of four 64-environment minibatches).  A fixture, not the oracle: nothing under oracle/ runs
    here.  Raises if the losses differ by more than the north-star 1e-4."""
    from a2c_ppo_acktr import algo
    from a2c_ppo_acktr.model import Policy
    from a2c_ppo_acktr.storage import RolloutStorage
    path = os.path.join(ROOT, "tests", "golden", "update_C5m.npz")
    if not os.path.exists(path):
        return {"checked": False, "why": "tests/golden/update_C5m.npz missing"}
    with np.load(path, allow_pickle=False) as z:
        case = {k: z[k] for k in z.files}
    g = json.loads(str(case["config"]))
    T, N = g["T"], g["N"]
    if obs_np is None:             # the bench itself runs on ring frames: regenerate the golden's independent stacks
        obs_np = synthetic.atari_frames(T + 1, N, 0, None, tuple(g["obs_shape"]))
    assert (T, N) == (cfg["T"], cfg["N"]) and abs(float(obs_np.astype(np.float64).sum()) - float(case["obs_sum"])) < 1e-3
    torch.manual_seed(1)
    pol = Policy(tuple(g["obs_shape"]), Discrete(6), base_kwargs={"recurrent": True}).to(dev)
    st = RolloutStorage(T, N, tuple(g["obs_shape"]), Discrete(6), 512, obs_dtype=torch.uint8)
    st.to(dev)
    rew, msk, bad = synthetic.atari_signals(T, N, 0)                      # frames: obs_np, the bench's own rollout
    st.obs.copy_(torch.from_numpy(obs_np))
    for k, a in (("rewards", rew), ("masks", msk), ("bad_masks", bad), ("actions", case["ro_actions"]),
                 ("action_log_probs", case["ro_action_log_probs"]), ("value_preds", case["ro_value_preds"])):
        getattr(st, k).copy_(torch.from_numpy(a))
    st.recurrent_hidden_states[0].copy_(torch.from_numpy(case["ro_h0"]))
    st.compute_returns(torch.from_numpy(case["ro_next_value"]).to(dev), g["use_gae"], g["gamma"], g["gae_lambda"],
                       g["use_proper_time_limits"])
    assert np.array_equal(st.returns.cpu().numpy()[:T], case["returns"][:T]), "returns differ from the reference"
    agent = algo.PPO(pol, g["clip_param"], g["ppo_epoch"], g["num_mini_batch"], g["value_loss_coef"], g["entropy_coef"],
                     lr=g["lr"], eps=g["eps"], max_grad_norm=g["max_grad_norm"])
    torch.manual_seed(7)
    got = np.array(agent.update(st))
    want = case["losses"]
    err = float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 2e-2)))
    res = {"checked": True, "case": "update_C5m (reference-generated golden, T=128, N=256, 4 x 8192-row recurrent minibatches)",
           "losses": got.tolist(), "reference_losses": want.tolist(), "max_rel_err": err, "tolerance": 1e-4}
    if not np.allclose(got, want, rtol=1e-4, atol=2e-6):
        raise SystemExit("bench.py: first-update losses differ from the reference golden: %s" % json.dumps(res))
    return res


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of the C5 shard
# (profiles/r02y_tcx_gru_full_c5.txt); only meaningful for that workload
NCU_DRAM_SOURCE = "profiles/r02y_tcx_gru_full_c5.txt (ncu --set full, C5 shard)"
NCU_DRAM_BYTES = {
    "conv1.fwd": 0.463882e9 + 784.483e6, "conv2.fwd": 0.841301e9 + 310.626e6, "conv3.fwd": 0.340044e9 + 82.657e6,
    "conv2.dgrad": 0.759749e9 + 794.737e6, "conv3.dgrad": 0.272990e9 + 291.951e6,
    "conv1.wgrad": 1.764634e9 + 3.855e6, "conv2.wgrad": 1.182154e9 + 7.643e6, "conv3.wgrad": 0.444832e9 + 9.184e6,
    "gru.seq_fwd": 0.053998e9 + 47.039e6, "gru.seq_bwd": 0.104061e9 + 62.479e6,
}


def describe(cfg):
    if cfg["algo"] == "ppo":
        return ("PPO %s, %d steps x %d procs per GPU, %d epochs x %d minibatches, clip %.2g%s"
                % ("CNNBase" if len(cfg["obs_shape"]) == 3 else "MLPBase", cfg["T"], cfg["N"], cfg["ppo_epoch"],
                   cfg["num_mini_batch"], cfg["clip_param"], ", GRU" if cfg["recurrent"] else ""))
    return "%s %s, %d steps x %d procs" % (cfg["algo"].upper(), "CNNBase", cfg["T"], cfg["N"])


# ---------------------------------------------------------------------------
# per-kernel timing (CUDA events around every op of one minibatch program)
# ---------------------------------------------------------------------------
def kernel_roofline(agent, st, cfg, dev, pk, measure_peak=None, reps=8):
    from a2c_ppo_acktr import _engine
    prog = getattr(agent, "_prog", None)
    if isinstance(prog, dict):           # ACKTR keeps three programs; profile the main fwd/bwd pass
        prog = prog.get("main")
    if prog is None:
        return None, None
    T, N = cfg["T"], cfg["N"]
    mbs = prog.builder.B if hasattr(prog, "builder") else T * N
    kind_names = {2: "reduce_splits", 3: "loss_epilogue", 4: "grad_norm", 5: "optim_step", 7: "gru_init",
                  8: "gru_gates", 9: "gru_gates_bwd", 10: "bias_pixel_sum", 11: "weight_prep", 12: "gru.seq_fwd",
                  13: "gru.seq_bwd", 16: "lo_plane", 17: "weight_image", 18: "frames_s2d", 19: "bias_colsum", 20: "heads.wgrad", 21: "weight_image", 24: "heads.fwd", 25: "heads.dgrad"}
    labels, flops = [], []
    plan_names = [p for p in getattr(prog, "plan_names", [])]
    gi = 0
    for (kind, flags, desc), (lab, fl) in zip(prog.entries, prog.labels):
        if kind == _engine.OP_GEMM:
            labels.append(plan_names[gi] if gi < len(plan_names) else "gemm")
            gi += 1
            flops.append(2.0 * desc.M * desc.N * desc.K * desc.batch)
        elif kind in (14, 15):              # TMA-fed tensor-core engine: labelled by layer
            labels.append(lab)
            flops.append(fl)
        elif kind in (22, 23):              # fused MLPBase: 19.7 kFLOP / row forward, 36.6 backward (d_in = 11)
            labels.append("mlp.fused_bwd" if kind == 23 else "mlp.fused_fwd")
            per_row = 2.0 * (2 * (desc.d_in * 64 + 64 * 64) + 64 * desc.zs)
            flops.append(desc.B * per_row * (2.0 if kind == 23 else 1.0))
        elif kind in (12, 13):              # persistent GRU recurrence: T steps of [N, H] x [H, 3H] (fp32 SIMT pipe)
            labels.append(kind_names[kind])
            flops.append(2.0 * desc.T * desc.N * 3 * desc.H * desc.H)
        else:
            labels.append(kind_names.get(kind, lab or "op%d" % kind))
            flops.append(0.0)
    singles = []
    for kind, flags, desc in prog.entries:
        p = _engine.Program(dev)
        p.add(kind, desc, flags)
        singles.append(p.finalize())
    if cfg.get("recurrent"):
        # a recurrent minibatch = E whole environments, rows time-major; initial state gathered per minibatch
        from a2c_ppo_acktr import _native as nat
        E = mbs // T
        rt = agent.actor_critic.runtime()
        env = torch.randperm(N)[:E].to(dev)
        idx = ((torch.arange(T, device=dev, dtype=torch.int64) * N).view(T, 1) + env.view(1, E)).reshape(-1).contiguous()
        hxs_mb = rt.arena.bufs["hxs_mb"][:E * rt.spec.hidden].view(E, rt.spec.hidden)
        nat.gather_rows([(st.recurrent_hidden_states[0], hxs_mb)], env, E)
    else:
        idx = torch.randperm(T * N)[:mbs].to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    times = np.zeros(len(singles))
    for r in range(reps + 2):
        flush.zero_()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(singles) + 1)]
        evs[0].record()
        for i, p in enumerate(singles):
            p.run(idx)
            evs[i + 1].record()
        torch.cuda.synchronize()
        if r >= 2:
            times += np.array([evs[i].elapsed_time(evs[i + 1]) for i in range(len(singles))])
    times /= reps
    total = times.sum()
    # aggregate by op label (a recurrent program repeats the per-time-step ops T times)
    agg = {}
    for lab, t, f in zip(labels, times, flops):
        a = agg.setdefault(lab, [0.0, 0.0, 0])
        a[0] += t; a[1] += f; a[2] += 1
    rows = [{"op": lab, "launches": a[2], "ms": float(a[0]), "share": float(a[0] / total),
             "tflops": float(a[1] / a[0] / 1e9) if a[1] else None} for lab, a in agg.items()]
    rows.sort(key=lambda r: -r["ms"])
    flop_rows = [r for r in rows if r["tflops"] is not None]
    if not flop_rows:
        return None, rows[:24]
    if measure_peak is not None:
        pk.update(measure_peak())
    from a2c_ppo_acktr import _native as nat
    tcx_ops = any(kind in (14, 15) for kind, _, _ in prog.entries)
    is_c5 = bool(cfg.get("recurrent")) and cfg["T"] == 128 and cfg["N"] == 256
    bf16_peak = pk["tensor_sustained"]
    tf32_peak = pk.get("tf32_tflops_sustained") or bf16_peak / 2.0
    tf32_src = ("measured in this run: " + pk["how"]) if pk.get("tf32_tflops_sustained") else "bf16 sustained / 2 (stand-in)"

    def describe_row(r):
        """Roofline entry of one op row: kernel name, the tensor-pipe peak of the arithmetic it issues, and what its
        fp32-parity scheme can reach of it (every fp32 product costs 3 MMAs; conv1.wgrad 2: uint8 frames are exact)."""
        op = r["op"]
        if op.startswith("gru.seq"):
            kern = "a2cb::gru_%s_%s_kernel" % (nat.gru_seq_variant(), "bwd" if op.endswith("bwd") else "fwd")
            f16 = nat.gru_seq_variant() == "tc"
            peak, mmas, what = (bf16_peak, 3, "kind::f16, fp16 hi/lo operand pairs") if f16 else (tf32_peak, 3, "mma.sync 3xTF32")
            bound = "sequential: T dependent steps, one cross-CTA exchange per step (latency), " + what
        elif op.startswith("mlp.fused"):
            kern, peak, mmas, bound = "a2cb::mlp_%s_kernel" % op[-3:], 148 * 128 * 2 * 1.965e-3, 1, "fp32 SIMT, latency-bound at 64 rows"
        elif tcx_ops and op == "conv1.fwd":
            kern, peak, mmas, bound = "a2cb::tcx_fwd_kernel (conv1.fwd, bf16 mode)", bf16_peak, 3, "tensor (3 x kind::f16 on exact bf16 frames)"
        elif tcx_ops:
            kern = "a2cb::tcx_%s_kernel (%s)" % ("wgrad" if op.endswith("wgrad") else "fwd", op)
            peak, mmas, bound = tf32_peak, (2 if op == "conv1.wgrad" else 3), "tensor (3xTF32 on tcgen05)"
        else:
            kern, peak, mmas, bound = "a2cb::gemm_tc_kernel / gemm_kernel (%s)" % op, tf32_peak, 3, "tensor (3xTF32) / fp32 SIMT"
        return {"kernel": kern, "op": op, "bound": "tensor", "bound_detail": bound, "achieved": r["tflops"], "peak": peak,
                "unit": "TFLOP/s", "frac": r["tflops"] / peak, "mmas_per_product": mmas,
                "frac_of_attainable": r["tflops"] / (peak / mmas), "share_of_step": r["share"],
                "launch_ms": r["ms"] / r["launches"], "launches_per_step": r["launches"],
                "traffic": NCU_DRAM_BYTES.get(op) if is_c5 else None}

    top = flop_rows[0]                                  # the largest-share launch of the minibatch program
    roof = describe_row(top)
    roof.update({
        "peak_source": "%s bf16_tflops_sustained = %.1f; TF32 dense peak %.1f TFLOP/s (%s).  `achieved` counts ALGORITHMIC fp32 "
                       "flops; `frac_of_attainable` divides by peak / mmas_per_product" % (pk["source"], bf16_peak, tf32_peak, tf32_src),
        "step_program_ms": float(total), "gemm_engine_mode": int(nat.lib().a2cb_get_gemm_mode()),
        "traffic_source": NCU_DRAM_SOURCE if (is_c5 and NCU_DRAM_BYTES.get(top["op"])) else None})
    conv = [r for r in flop_rows if not r["op"].startswith("gru.seq") and r is not top]
    if conv:
        roof["next_kernel"] = describe_row(conv[0])      # the largest dense-layer launch besides the dominant one
    return roof, rows[:24]


def scan_roofline(dev, pk):
    """GAE scan HBM GB/s on a [128, 2^20] sweep (config-sized scans are launch-bound: 20 KB..5 MB)."""
    from a2c_ppo_acktr import _native
    T, N = 128, 1 << 20
    r = torch.randn(T, N, 1, device=dev)
    v = torch.randn(T + 1, N, 1, device=dev)
    m = (torch.rand(T + 1, N, 1, device=dev) > 0.01).float()
    b = torch.ones(T + 1, N, 1, device=dev)
    out = torch.zeros(T + 1, N, 1, device=dev)
    nv = torch.randn(N, device=dev)
    ts = []
    for i in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _native.returns_scan(r, v, out, m, b, nv, 0.99, 0.95, True, True, 1)
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    ms = float(np.mean(ts))
    gbs = 20.0 * T * N / ms / 1e6
    return {"kernel": "returns_scan_env_kernel<GAE,PTL,4,4>", "shape": [T, N], "bytes_per_transition": 20,
            "bound": "hbm", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": gbs / pk["hbm"], "ms": ms,
            "note": "inputs 2.7 GB > L2"}


# ---------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle port of the reference on the host cores
# ---------------------------------------------------------------------------
def cpu_update_once(cfg, state):
    from oracle import returns as o_ret
    from oracle import update as o_upd
    ro, p, spec = state["ro"], state["p"], state["spec"]
    t0 = time.perf_counter()
    ret, vp = o_ret.compute_returns(ro["rewards"].numpy(), ro["value_preds"].numpy(), ro["masks"].numpy(),
                                    ro["bad_masks"].numpy(), ro["next_value"].numpy(), cfg["use_gae"], cfg["gamma"],
                                    cfg["gae_lambda"], cfg["use_proper_time_limits"])
    ro["returns"] = torch.from_numpy(ret)
    ro["value_preds"] = torch.from_numpy(vp)
    if cfg["algo"] == "ppo":
        out = o_upd.ppo_update(p, spec, ro, cfg, optimizer=state.get("opt"))
    elif cfg["algo"] == "acktr":
        from oracle import kfac as o_kfac
        if "opt" not in state:
            state["opt"] = o_kfac.KFACState()
        out = o_kfac.acktr_update(p, ro, cfg, state["opt"]) + (state["opt"],)
    else:
        out = o_upd.a2c_update(p, spec, ro, cfg, optimizer=state.get("opt"))
    state["opt"] = out[3]
    for k in ("obs", "recurrent_hidden_states", "masks", "bad_masks"):       # after_update
        ro[k][0].copy_(ro[k][-1])
    return time.perf_counter() - t0, out[:3]


def cpu_sample_config(cfg, envs):
    """The same workload on `envs` environments (the per-transition cost of the reference path does not depend on it)."""
    c = dict(cfg)
    c["N"] = max(int(envs), c.get("num_mini_batch", 1))
    return c


def cpu_state(cfg, inputs=None):
    from oracle import policy as o_pol
    from oracle import update as o_upd
    spec = o_pol.Spec(cfg["obs_shape"], cfg["action"], cfg["recurrent"])
    p0 = o_pol.init_params(spec, seed=1)
    torch.manual_seed(2)
    if inputs is not None:
        obs, rew, msk, bad = inputs
        inputs = (obs.astype(np.float32) if obs.dtype != np.float32 else obs, rew, msk, bad)
    with torch.no_grad():
        ro = synthetic.make_rollout(cfg, lambda o, h, m: o_pol.act(p0, spec, o, h, m),
                                    lambda o, h, m: o_pol.get_value(p0, spec, o, h, m), spec.hidden_state_size, seed=0,
                                    inputs=inputs)
    return {"ro": ro, "p": o_upd.make_leaf_params(p0), "spec": spec}


def cpu_time_updates(cfg, threads, steps, warmup, inputs=None, budget_s=None):
    """Times `steps` updates (after `warmup`) of the oracle port on `threads` host threads; stops early when the
    timed updates have used `budget_s` seconds.  -> (seconds per update, updates timed, last losses)."""
    torch.set_num_threads(threads)
    state = cpu_state(cfg, inputs)
    ts, losses, used = [], None, 0.0
    for i in range(warmup + steps):
        torch.manual_seed(7 + i)
        dt, losses = cpu_update_once(cfg, state)
        if i >= warmup:
            ts.append(dt)
            used += dt
            if budget_s is not None and used >= budget_s:
                break
    return float(np.mean(ts)), len(ts), losses


def cpu_baseline(cfg, inputs=None):
    """Reported baseline (not a target): the oracle port of the reference on this box's host cores, on a bounded
    sample of the SAME workload -- ONE full-size update with all cores (no warm-up: a C5 update is ~20 s), plus the
    reference's own training setting `torch.set_num_threads(1)` (reference main.py:38) on a reduced number of
    environments."""
    cores = os.cpu_count() or 1
    big = cfg["T"] * cfg["N"] > 8192
    sec, n, _ = cpu_time_updates(cfg, cores, 1 if big else 3, 0 if big else 1, inputs)
    tr = cfg["T"] * cfg["N"]
    out = {"value": tr / sec, "unit": "transitions/s", "cores": cores, "kind": "port",
           "sample": "%d full-size %s update%s (%d transitions each, %s) of the oracle port (torch %s CPU, %d threads), "
                     "%.3f s/update" % (n, cfg["name"], "s" if n > 1 else "", tr,
                                        "no warm-up" if big else "1 warm-up", torch.__version__, cores, sec)}
    c1 = cpu_sample_config(cfg, max(cfg.get("num_mini_batch", 1) * 4, 2048 // cfg["T"])) if big else cfg
    sec1, n1, _ = cpu_time_updates(c1, 1, 1, 0 if big else 1)
    tr1 = c1["T"] * c1["N"]
    out["threads_1"] = {"value": tr1 / sec1, "unit": "transitions/s", "cores": 1,
                        "sample": "%d update of %d transitions (%d of %d environments) with torch.set_num_threads(1), the "
                                  "reference's own training setting (main.py:38), %.3f s/update" % (n1, tr1, c1["N"], cfg["N"], sec1)}
    torch.set_num_threads(cores)
    return out


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path (its oracle port: the pure-Python reference cannot
    travel to the GPU box) on ALL host cores, on the ours-arm's config at FULL size (every environment of one GPU's
    shard).  One warm-up update; the timed updates stop after A2CB_REF_BUDGET_S seconds (default 360) so that the run
    ends within minutes whatever --steps says -- the sample string says how many were timed."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = workload_config(args.workload)
    cores = os.cpu_count() or 1
    budget = float(os.environ.get("A2CB_REF_BUDGET_S", "360"))
    big = cfg["T"] * cfg["N"] > 8192
    warm = min(args.warmup, 1) if big else args.warmup
    sec, n, losses = cpu_time_updates(cfg, cores, args.steps, warm, None, budget)
    tr = cfg["T"] * cfg["N"]
    val = tr / sec
    sample = ("%d of the %d requested full-size %s updates (%d transitions each, %d warm-up; time budget %.0f s) of the "
              "oracle port of the reference (torch %s CPU, %d threads), %.3f s/update"
              % (n, args.steps, cfg["name"], tr, warm, budget, torch.__version__, cores, sec))
    print(json.dumps({
        "impl": "reference", "metric": "%s update transitions/sec" % cfg["algo"].upper(),
        "value": val, "unit": "transitions/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "steps_timed": n, "ms_per_step": 1e3 * sec, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s: %s" % (cfg["name"], describe(cfg)), "envs_per_gpu": cfg["N"], "num_steps": cfg["T"],
                   "obs": "fp32 (reference storage)", "step": "compute_returns + update + after_update"},
        "cpu_baseline": {"value": val, "unit": "transitions/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "transitions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "losses": [float(x) for x in losses],
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C5", choices=sorted(synthetic.CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-numa-bind", action="store_true",
                    help="do not pin the process to the CPU cores of the GPU's NUMA node before the pinned rollout is allocated")
    ap.add_argument("--no-parity", action="store_true", help="skip the first-update check against the reference golden (C5)")
    ap.add_argument("--stacked-frames", action="store_true",
                    help="Atari workloads: store / upload full 4-frame stacks (reference layout) instead of the frame ring")
    ap.add_argument("--envs", type=int, default=0, help="override the number of environments per GPU")
    ap.add_argument("--recurrent-sharding", default="", choices=["", "global", "balanced", "stratified"],
                    help="N > 1, recurrent PPO: 'global' / 'balanced' = the reference's one randperm(N_total), every rank "
                         "processing the environments it owns / exactly E / world of them after migrating the surplus; "
                         "'stratified' = balanced per-rank permutations (not the reference's minibatch composition)")
    args = ap.parse_args()
    global ENVS_OVERRIDE
    ENVS_OVERRIDE = args.envs
    # stdout carries exactly ONE JSON line: everything else any library writes to file descriptor 1 during the run (NCCL
    # prints its version banner there) is sent to stderr; the JSON line goes to the saved descriptor at the end.
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    real_stdout = sys.stdout
    sys.stdout = os.fdopen(saved, "w", buffering=1)
    try:
        if args.impl == "reference":
            run_reference(args)
        else:
            if args.warmup < 3:
                args.warmup = 3
            run_ours(args)
    finally:
        sys.stdout.flush()
        sys.stdout = real_stdout


if __name__ == "__main__":
    main()
