"""profiling helper: fused step + observation planes (pom_batch_step_observe) and the standalone observe kernel"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pomcpp_b200 as pb
n = 1 << 20
b = pb.Batch(n, n_templates=4096, max_ticks=800)
b.rollout(96, 5, 0, 0)
stride = int(pb.lib().pom_batch_obs_stride(b.h))
obs = b.alloc(stride * pb.OBS_BYTES)
mv = b.alloc(4 * n)
b.generate_moves(mv, 1, 0, 6)
flags = pb.STEP_AUTORESET | pb.STEP_COUNT
for name, fn in (("step", lambda: b.step(mv, flags)), ("step_observe(1 agent)", lambda: b.step_observe(mv, obs, 1, 4, flags)),
                 ("observe_planes(1 agent)", lambda: pb._ck(pb.lib().pom_batch_observe_planes(b.h, obs, 1, 4)))):
    for _ in range(3):
        fn()
    b.sync()
    b.event(0)
    for _ in range(20):
        fn()
    b.event(1)
    print("%-26s %.4f ms per 1 Mi envs" % (name, b.elapsed_ms() / 20))
