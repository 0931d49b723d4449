"""profiling helper: the fused rollout with random agents (configs[3] shape, shorter): steady-state mix, then one launch to capture"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pomcpp_b200 as pb
n = int(os.environ.get("POM_PROF_ENVS", 1 << 19))
b = pb.Batch(n, n_templates=4096, max_ticks=800)
b.rollout(96, 5, 0, 0)
b.sync()
b.clear_stats()
b.event(0)
b.rollout(200, 5, 96, 0)
b.event(1)
ms = b.elapsed_ms()
print("rollout: %d envs x 200 ticks in %.3f ms = %.3e env-steps/s" % (n, ms, n * 200 / ms * 1e3))
