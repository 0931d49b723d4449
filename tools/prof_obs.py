"""profiling helper: fused step + observation planes (pom_batch_step_observe) and the standalone observe kernel, on the
bench's workload (a ring of different random move sets: with ONE move set repeated every tick the games degenerate - few
bombs and flames - and the observation kernel looks twice as fast as it is)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pomcpp_b200 as pb
n = int(os.environ.get("POM_PROF_ENVS", 1 << 20))
b = pb.Batch(n, n_templates=4096, max_ticks=800)
b.rollout(96, 5, 0, 0)
stride = int(pb.lib().pom_batch_obs_stride(b.h))
obs = b.alloc(stride * pb.OBS_BYTES)
RING = 23
mv = b.alloc(4 * n * RING).value
for k in range(RING):
    b.generate_moves(mv + 4 * n * k, 1, k, 6)
flags = pb.STEP_AUTORESET | pb.STEP_COUNT
tick = [0]
def moves():
    tick[0] += 1
    return mv + 4 * n * (tick[0] % RING)
for name, fn in (("step", lambda: b.step(moves(), flags)), ("step_observe(1 agent)", lambda: b.step_observe(moves(), obs, 1, 4, flags)),
                 ("observe_planes(1 agent)", lambda: pb._ck(pb.lib().pom_batch_observe_planes(b.h, obs, 1, 4))),
                 ("observe_planes_cropped(1 agent)", lambda: pb._ck(pb.lib().pom_batch_observe_planes_cropped(b.h, obs, 1, 4)))):
    for _ in range(30):
        fn()
    b.sync()
    b.event(0)
    for _ in range(20):
        fn()
    b.event(1)
    print("%-26s %.4f ms per launch of %d envs" % (name, b.elapsed_ms() / 20, n))
