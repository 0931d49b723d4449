"""profiling helper: configs[4] tree-search expansion (pom_batch_expand_step), kernel time by CUDA events"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pomcpp_b200 as pb
n = 1 << 20
b = pb.Batch(n, n_templates=4096, max_ticks=800)
b.rollout(96, 5, 0, 0)
ROOTS, FAN = 4096, 1296
eb = pb.Batch(ROOTS * FAN, n_templates=16, empty=True)
roots = (np.arange(ROOTS, dtype=np.uint32) * 251) % n
for _ in range(2):
    eb.expand_step_from(b, roots, FAN, 0)
eb.sync()
eb.event(0)
for _ in range(10):
    eb.expand_step_from(b, roots, FAN, 0)
eb.event(1)
ms = eb.elapsed_ms() / 10
print("expand: %.4f ms per %d children (events, 10 back to back) = %.3e children/s, %.0f GB/s written" %
      (ms, ROOTS * FAN, ROOTS * FAN / ms * 1e3, ROOTS * FAN * 292 / ms / 1e6))
t0 = time.perf_counter()
for _ in range(10):
    eb.expand_step_from(b, roots, FAN, 0)
    eb.sync()
print("expand: %.4f ms wall per synchronous call" % ((time.perf_counter() - t0) * 100))
