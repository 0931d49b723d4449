#!/usr/bin/env python
"""bench.py — env-steps/sec of the batched Pommerman step path on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Own arm (default): BASELINE.json configs[2] — 1,048,576 envs per GPU, random joint actions incl. bombs,
kicks and chain explosions, PER-TICK kernel (pom_batch_step, auto-reset on so every env is live on every
tick).  One "step" = one tick over the whole batch = n_envs env-steps per GPU.  Per-tick joint actions are
pre-generated ON the device for a ring of ticks (inputs resident in HBM when the timed region starts);
the 306 MB record array per GPU is larger than the 126 MB L2, so no flush is needed between steps.
K steps are timed with CUDA events on the launching stream, bracketed by a barrier + device sync, MAX
over ranks.  `e2e` repeats the measurement through pom_batch_step_host with pinned HOST buffers (moves in,
status bytes out, every step).  `cpu_baseline` times the UNMODIFIED reference (oracle/_ref) on the box's
host cores on a bounded sample of the same workload (rank 0, N=1 only).

Reference arm (--impl reference): the reference's own CPU Step on all host cores, one step = one tick over
a bounded sample of envs; same metric / unit / config.

Envs are independent: they shard over GPUs with no data-path collective ("scaling": "weak", per-GPU work
fixed); the only collective is the final NCCL all-reduce of the episode counters.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENVS_PER_GPU = 1 << 20
N_TEMPLATES = 4096
MOVE_RING = 256              # distinct pre-generated ticks of joint actions, cycled
PREROLL_TICKS = 96           # untimed fused rollout so the batch is in its steady-state mix
RNG_SEED = 20240229
ALGO_BYTES = 2 * 289 + 4     # packed state in + out + 4 move bytes per env-step
CPU_SAMPLE_ENVS = 262144
CPU_SAMPLE_TICKS = 96
E2E_STEPS = 50


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU with NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.stop_flag = False
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                 "sw_thermal_slowdown": 0x20, "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.01)

    def result(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def load_traffic():
    """dram bytes per launch of k_step from the committed ncu capture (profiles/), or None."""
    p = os.path.join(ROOT, "profiles", "k_step_ncu_summary.json")
    try:
        return json.load(open(p)).get("dram_bytes_per_launch")
    except Exception:
        return None


def cpu_reference_states(n, preroll):
    """Initial states for the CPU legs, built with the reference's own InitState on clean seeds and
    pre-rolled with the reference's Step so the sample is in the steady-state mix."""
    import oracle
    oracle.build()
    R = oracle.reference() if oracle.have_reference() else None
    O = oracle.restatement()
    eng = R if R is not None else O
    kind = "reference" if R is not None else "port"
    seeds = oracle.clean_seeds(N_TEMPLATES)
    T = eng.zero_state(N_TEMPLATES)
    for k, sd in enumerate(seeds):
        eng.init_state(T[k:k + 1], sd)
    S = T[np.arange(n) % N_TEMPLATES].copy()
    status = np.zeros(n, np.uint8)
    cores = os.cpu_count() or 1
    if preroll:
        mv = np.stack([O.rng_moves(RNG_SEED, 0, n, t, 6) for t in range(preroll)])
        eng.bench_steps(S, status, mv, cores, T)
    return eng, O, kind, S, status, T, cores


def cpu_baseline_leg():
    eng, O, kind, S, status, T, cores = cpu_reference_states(CPU_SAMPLE_ENVS, 32)
    mv = np.stack([O.rng_moves(RNG_SEED, 0, CPU_SAMPLE_ENVS, 1000 + t, 6) for t in range(CPU_SAMPLE_TICKS)])
    best = 0.0
    for _ in range(2):
        t, steps = eng.bench_steps(S, status, mv, cores, T)
        best = max(best, steps / t)
    return {"value": best, "unit": "env-steps/s", "cores": cores, "kind": kind,
            "sample": "%d envs x %d ticks (auto-reset), best of 2, reference Step -O3, %d threads" %
                      (CPU_SAMPLE_ENVS, CPU_SAMPLE_TICKS, cores)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = CPU_SAMPLE_ENVS
    eng, O, kind, S, status, T, cores = cpu_reference_states(n, 32)
    ring = np.stack([O.rng_moves(RNG_SEED, 0, n, 2000 + t, 6) for t in range(64)])
    for w in range(args.warmup):
        eng.bench_steps(S, status, ring[w % 64:w % 64 + 1], cores, T)
    total_t, total_steps = 0.0, 0
    for k in range(args.steps):
        t, steps = eng.bench_steps(S, status, ring[k % 64:k % 64 + 1], cores, T)
        total_t += t
        total_steps += steps
    v = total_steps / total_t
    line = {"impl": "reference", "metric": "env-steps/sec", "value": v, "unit": "env-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total_t / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": "configs[2]: random joint actions incl. bombs/kicks/chains, reference bboard::Step "
                                   "on host cores; one step = one tick over a bounded sample of %d envs" % n,
                       "envs_per_step": n, "threads": cores},
            "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": cores, "kind": kind,
                             "sample": "%d envs per step, %d steps" % (n, args.steps)},
            "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_own(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    torch = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import pomcpp_b200 as pb
    from pomcpp_b200 import shard
    n = ENVS_PER_GPU
    K, W = args.steps, args.warmup
    plan = shard.shard_plan(rank, world, n)
    b = pb.Batch(n, device=local_rank, env_offset=plan["first"], n_templates=N_TEMPLATES, max_ticks=800)
    ring = min(MOVE_RING, K + W)
    moves_dev = b.alloc(4 * n * ring)
    for t in range(ring):
        b.generate_moves(moves_dev.value + 4 * n * t, RNG_SEED, 100000 + t, 6)
    b.rollout(PREROLL_TICKS, RNG_SEED, 0, 0)          # untimed: reach the steady-state mix
    b.sync()
    flags = pb.STEP_AUTORESET | pb.STEP_COUNT

    # ---- the kernel alone: one launch per tick over the whole batch, ticks strictly one after the other
    SINGLE_STEPS = min(200, K)
    for w in range(5):
        b.step(moves_dev.value + 4 * n * (w % ring), flags)
    b.sync()
    b.event(0)
    for k in range(SINGLE_STEPS):
        b.step(moves_dev.value + 4 * n * (k % ring), flags)
    b.event(1)
    single_ms = b.elapsed_ms() / SINGLE_STEPS
    b.sync()
    b.clear_stats()
    # ---- the timed region proper: the same kernel, the two halves of the batch on two streams (POM_STEP_OVERLAP), so
    #      that the partly filled last wave of one launch runs next to the first wave of the next
    if os.environ.get("POM_BENCH_OVERLAP", "1") != "0":
        flags |= pb.STEP_OVERLAP

    def barrier():
        b.sync()
        if dist is not None:
            torch.cuda.synchronize()
            dist.barrier()

    for w in range(W):
        b.step(moves_dev.value + 4 * n * (w % ring), flags)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = b.launch_count()
    b.event(0)
    for k in range(K):
        b.step(moves_dev.value + 4 * n * ((W + k) % ring), flags)
    b.event(1)
    ms = b.elapsed_ms()
    barrier()
    sampler.stop_flag = True
    sampler.join()
    launches = b.launch_count() - l0
    stats = b.stats()
    assert stats.env_steps == n * (K + W), "kernel did not step every env on every tick"

    # ---- e2e: host buffers through pom_batch_step_host, copies inside the timed region
    #      (a) one batch, pom_batch_step_host: launch, wait, launch, ... ; (b) the same envs as two half-batches stepped
    #      alternately through pom_batch_step_host_async + pom_batch_sync, the way an actor loop with two env groups
    #      runs: while the host consumes the results of one half, the GPU steps the other.  In both, every tick moves
    #      all 4 move bytes/env host -> device and the status byte/env device -> host, and the host waits for them.
    E2E_RING = 32                                        # distinct move sets: a short ring would make the games periodic
    mv_ring = [pb.pinned_array((n, 4), np.uint8) for _ in range(E2E_RING)]
    st_host, st_owner = pb.pinned_array((n,), np.uint8)
    rng = np.random.default_rng(RNG_SEED + rank)
    for arr, _ in mv_ring:
        arr[:] = rng.integers(0, 6, size=(n, 4), dtype=np.uint8)      # the policy's output, already in pinned memory
    e2e_flags = flags & ~pb.STEP_OVERLAP               # every tick is waited for
    for w in range(8):
        b.step_host(mv_ring[w % E2E_RING][0], st_host, e2e_flags)
    barrier()
    t0 = time.perf_counter()
    for k in range(E2E_STEPS):
        b.step_host(mv_ring[k % E2E_RING][0], st_host, e2e_flags)
    e2e_single_s = time.perf_counter() - t0
    barrier()

    h = n // 2
    halves = [pb.Batch(h, device=local_rank, env_offset=plan["first"] + i * h, n_templates=N_TEMPLATES, max_ticks=800) for i in range(2)]
    for x in halves:
        x.rollout(PREROLL_TICKS, RNG_SEED, 0, 0)
        x.sync()
    mv_half = [[arr[i * h:(i + 1) * h] for arr, _ in mv_ring] for i in range(2)]
    st_half = [st_host[i * h:(i + 1) * h] for i in range(2)]

    def double_buffered(steps):
        halves[0].step_host_async(mv_half[0][0], st_half[0], e2e_flags)
        for k in range(steps):
            halves[1].step_host_async(mv_half[1][k % E2E_RING], st_half[1], e2e_flags)
            halves[0].sync()                   # results of half 0 are on the host now; its next moves are written here
            if k + 1 < steps:
                halves[0].step_host_async(mv_half[0][(k + 1) % E2E_RING], st_half[0], e2e_flags)
            halves[1].sync()                   # results of half 1

    double_buffered(8)
    barrier()
    t0 = time.perf_counter()
    double_buffered(E2E_STEPS)
    e2e_s = time.perf_counter() - t0
    barrier()
    e2e_launches = halves[0].launch_count() + halves[1].launch_count()
    for x in halves:
        x.close()

    # ---- aggregate over ranks
    dev = "cuda" if dist is not None else "cpu"
    ms_max, e2e_max, e2e_single_max = [float(v) for v in shard.max_over_ranks([ms, e2e_s, e2e_single_s], dist, dev)]
    # the one collective of the run: final NCCL reduce of the episode counters
    counters = shard.reduce_counters(b.stats().as_array(), dist, dev)

    if rank == 0:
        peak, peak_src = measured_peak()
        value = world * n * K / (ms_max * 1e-3)
        ms_per_step = ms_max / K
        achieved = ALGO_BYTES * n / (ms / K * 1e-3) / 1e9           # this rank's kernel, GB/s
        line = {
            "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32", "data": "synthetic",
            "config": {"workload": "configs[2]: %d envs per GPU x 4 random agents (uniform{0..5}: moves, bombs, kicks, chain "
                                   "explosions), per-tick kernel pom_batch_step, auto-reset%s" %
                                   (n, ", POM_STEP_OVERLAP (two half-batch launches per tick on two streams)" if flags & pb.STEP_OVERLAP else ""),
                       "envs_per_gpu": n, "envs_total": world * n, "record_bytes": 292, "templates": N_TEMPLATES,
                       "preroll_ticks": PREROLL_TICKS, "move_ring_ticks": ring,
                       "l2": "record array %d MB per GPU > 126 MB L2: every step streams from HBM, no flush needed" % (n * 292 // 2 ** 20),
                       "parallelism": "env-sharded x%d, no data-path collective" % world},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": load_traffic(), "algorithmic_bytes_per_env_step": ALGO_BYTES,
                         "peak_source": peak_src, "kernel": "k_step<128>", "step_ms": ms / K,
                         "launches_per_step": 2 if flags & pb.STEP_OVERLAP else 1,
                         "note": "achieved = 582 B x envs / time per tick over the timed region (CUDA events bracket both streams); traffic = DRAM bytes "
                                 "per tick from the ncu capture of one whole-batch launch (the two half-batch launches move the same bytes)",
                         "single_launch": {"launch_ms": single_ms, "achieved": ALGO_BYTES * n / (single_ms * 1e-3) / 1e9,
                                           "frac": ALGO_BYTES * n / (single_ms * 1e-3) / 1e9 / peak, "steps": SINGLE_STEPS,
                                           "what": "one launch per tick over the whole batch, no overlap between ticks"}},
            "e2e": {"value": world * n * E2E_STEPS / e2e_max, "unit": "env-steps/s",
                    "h2d_bytes_per_step": 4 * n, "d2h_bytes_per_step": n, "steps": E2E_STEPS,
                    "api": "two half-batches stepped alternately with pom_batch_step_host_async + pom_batch_sync (pinned host moves in, "
                           "status bytes out, the host waits for every half-batch every tick; the step kernel reads the moves from and "
                           "writes the status bytes to the pinned host buffers over PCIe itself)",
                    "single_batch": {"value": world * n * E2E_STEPS / e2e_single_max, "unit": "env-steps/s",
                                     "api": "one batch, pom_batch_step_host: launch and wait, tick after tick"}},
            "gpu_launches": int(launches),
            "clocks": sampler.result(),
            "episode_stats": {"env_steps": int(counters[0]), "episodes": int(counters[1]),
                              "wins": [int(x) for x in counters[2:6]], "draws": int(counters[6]),
                              "truncated": int(counters[7]), "sum_episode_len": int(counters[8]),
                              "invalid": int(counters[9]), "reduced_with": "nccl all_reduce" if world > 1 else "single rank"},
        }
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline_leg()
        print(json.dumps(line), flush=True)

    for _, owner in mv_ring:
        pb.pinned_free(owner)
    pb.pinned_free(st_owner)
    b.free(moves_dev)
    b.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
