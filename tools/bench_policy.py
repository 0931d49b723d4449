"""Throughput of the device-side SimpleAgent policy on one GPU; prints one JSON line per mode.

    python tools/bench_policy.py [--envs N] [--ticks K]

 rollout_simple4   fused rollout, four SimpleAgents per env (the reference's own benchmark setting,
                   unit_test/bboard/performance_test.cpp:38,59-63), auto-reset, 800-tick limit
 rollout_simple1v3 agent 0 uniform random, agents 1-3 SimpleAgent
 rollout_random    the headline fused rollout (no policy) for comparison
 pertick_simple4   pom_batch_policy_moves + pom_batch_step per tick (moves stay on the device)
 cpu_reference     the compiled reference (oracle/_ref): SimpleAgent::act x4 + bboard::Step, all host threads
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pomcpp_b200 as pb  # noqa: E402


def timed(b, fn, reps):
    fn()
    b.sync()
    b.clear_stats()
    b.event(0)
    for _ in range(reps):
        fn()
    b.event(1)
    b.sync()
    return b.elapsed_ms() / 1e3, b.stats().env_steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1 << 20)
    ap.add_argument("--ticks", type=int, default=200)
    ap.add_argument("--cpu-envs", type=int, default=4096)
    ap.add_argument("--cpu", action="store_true",
                    help="also time the compiled reference (its self-seeded SimpleAgents reach defect D5 - unbounded recursion - once in a "
                         "while and the process then dies of a stack overflow: opt-in)")
    a = ap.parse_args()
    b = pb.Batch(a.envs, n_templates=4096, max_ticks=800)
    tick = [0]

    def roll(flags):
        def f():
            b.rollout(a.ticks, 99, tick[0], flags)
            tick[0] += a.ticks
        return f

    for name, flags in (("rollout_random", 0), ("rollout_simple4", pb.ROLL_SIMPLE(15)), ("rollout_simple1v3", pb.ROLL_SIMPLE(14))):
        b.reset()
        tick[0] = 0
        sec, steps = timed(b, roll(flags), 3)
        st = b.stats()
        print(json.dumps({"mode": name, "env_steps_per_s": steps / sec, "envs": a.envs, "ticks": 3 * a.ticks,
                          "mean_episode_len": st.sum_episode_len / max(1, st.episodes), "episodes": st.episodes,
                          "draws": st.draws, "wins": list(st.wins), "truncated": st.truncated, "invalid": st.invalid}))
    b.reset()
    moves = b.alloc(4 * a.envs)
    tick[0] = 0

    def per_tick():
        for _ in range(20):
            b.policy_moves(moves, 99, tick[0], 15)
            b.step(moves, pb.STEP_AUTORESET | pb.STEP_COUNT)
            tick[0] += 1
    for _ in range(3):
        per_tick()            # let the games develop before timing
    sec, steps = timed(b, per_tick, 5)
    print(json.dumps({"mode": "pertick_simple4", "env_steps_per_s": steps / sec, "envs": a.envs, "ticks": 100}))
    b.free(moves)
    b.close()

    if not a.cpu:
        return
    try:
        import oracle
        R = oracle.reference()
    except Exception as e:                       # noqa: BLE001
        print(json.dumps({"mode": "cpu_reference", "unavailable": str(e)}))
        return
    import ctypes as C
    L = R.lib
    if not hasattr(L, "ref_bench_simple"):
        return
    O = oracle.restatement()
    n = a.cpu_envs
    seeds = oracle.clean_seeds(64)
    S = O.zero_state(n)
    for i in range(n):
        O.init_state(S[i:i + 1], seeds[i % 64])
    T = S[:64].copy()
    L.ref_bench_simple.restype = C.c_double
    L.ref_bench_simple.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_uint64, C.c_void_p]
    steps = C.c_ulonglong(0)
    thr = R.hardware_concurrency()
    t0 = time.time()
    sec = L.ref_bench_simple(S.ctypes.data, n, 400, thr, T.ctypes.data, 64, 99, C.byref(steps))
    print(json.dumps({"mode": "cpu_reference", "env_steps_per_s": steps.value / sec, "threads": thr, "envs": n, "ticks": 400,
                      "wall_s": time.time() - t0}))


if __name__ == "__main__":
    main()
