"""e2e (host moves in, status bytes out, sync per tick) of pom_batch_step_host for the current POM_CHUNKS setting."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pomcpp_b200 as pb

n = 1 << 20
b = pb.Batch(n, n_templates=4096, max_ticks=800)
b.rollout(96, 7, 0, 0)
mv, own1 = pb.pinned_array((16, n, 4), np.uint8)
st, own2 = pb.pinned_array((n,), np.uint8)
for t in range(16):
    mv[t] = pb.rng_moves(11, 0, n, 1000 + t, 6)
for t in range(20):
    b.step_host(mv[t % 16], st, pb.STEP_AUTORESET)
t0 = time.perf_counter()
K = 300
for t in range(K):
    b.step_host(mv[t % 16], st, pb.STEP_AUTORESET)
dt = (time.perf_counter() - t0) / K
print("POM_CHUNKS=%s  %.1f us/tick  %.3e env-steps/s" % (os.environ.get("POM_CHUNKS", "default"), dt * 1e6, n / dt))
