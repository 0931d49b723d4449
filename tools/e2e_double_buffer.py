"""e2e with two half-batches stepped alternately through pom_batch_step_host_async (host buffers in and out every tick,
one sync per half-batch and tick) against one whole batch through pom_batch_step_host."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pomcpp_b200 as pb

n = 1 << 20
K = 300
fl = pb.STEP_AUTORESET | pb.STEP_COUNT
# one batch, synchronous
b = pb.Batch(n, n_templates=4096, max_ticks=800)
b.rollout(96, 7, 0, 0)
R = 32
mv = [pb.pinned_array((n, 4), np.uint8)[0] for _ in range(R)]
st = pb.pinned_array((n,), np.uint8)[0]
for t in range(R):
    mv[t][:] = pb.rng_moves(11, 0, n, 1000 + t, 6)
for t in range(20):
    b.step_host(mv[t % R], st, fl)
t0 = time.perf_counter()
for t in range(K):
    b.step_host(mv[t % R], st, fl)
dt = (time.perf_counter() - t0) / K
print("one batch, step_host            %.1f us per 1 Mi env-steps  %.3e env-steps/s" % (dt * 1e6, n / dt))
b.close()
# two half-batches, double-buffered
h = n // 2
B = [pb.Batch(h, env_offset=i * h, n_templates=4096, max_ticks=800) for i in range(2)]
for x in B:
    x.rollout(96, 7, 0, 0)
sts = [pb.pinned_array((h,), np.uint8)[0] for _ in range(2)]
mvh = [[m[i * h:(i + 1) * h] for m in mv] for i in range(2)]
def run(steps):
    B[0].step_host_async(mvh[0][0], sts[0], fl)
    for t in range(steps):
        B[1].step_host_async(mvh[1][t % R], sts[1], fl)
        B[0].sync()                      # the host would consume sts[0] here and write the next moves of half 0
        if t + 1 < steps:
            B[0].step_host_async(mvh[0][(t + 1) % R], sts[0], fl)
        B[1].sync()                      # ... and sts[1] here
run(20)
t0 = time.perf_counter()
run(K)
dt = (time.perf_counter() - t0) / K
print("two half-batches, async + sync  %.1f us per 1 Mi env-steps  %.3e env-steps/s" % (dt * 1e6, n / dt))
