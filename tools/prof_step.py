"""profiling helper: per-tick kernel, one whole-batch launch per tick, by CUDA events; env POM_STEP_PINGPONG=0/1, POM_PROF_ENVS"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pomcpp_b200 as pb
n = int(os.environ.get("POM_PROF_ENVS", 1 << 20))
if os.environ.get("POM_PROF_HOLDER"):
    holder = pb.Batch(int(os.environ["POM_PROF_HOLDER"]), n_templates=16)      # another handle's allocations first
b = pb.Batch(n, n_templates=4096, max_ticks=800)
b.rollout(96, 5, 0, 0)
ring = 23
mv = b.alloc(4 * n * ring)
for k in range(ring):
    b.generate_moves(mv.value + 4 * n * k if hasattr(mv, "value") else mv + 4 * n * k, 1, k, 6)
base = mv.value if hasattr(mv, "value") else mv
flags = pb.STEP_AUTORESET | pb.STEP_COUNT
for name, fl in (("single launch per tick", flags), ("POM_STEP_OVERLAP", flags | pb.STEP_OVERLAP)):
    for k in range(20):
        b.step(base + 4 * n * (k % ring), fl)
    b.sync()
    for reps in (20, 200):
        b.event(0)
        for k in range(reps):
            b.step(base + 4 * n * (k % ring), fl)
        b.event(1)
        ms = b.elapsed_ms() / reps
        print("%-24s %4d ticks: %.4f ms per tick = %.3e env-steps/s, %.0f GB/s algorithmic (%.3f of 6546)" %
              (name, reps, ms, n / ms * 1e3, 582 * n / ms / 1e6, 582 * n / ms / 1e6 / 6546.2))
