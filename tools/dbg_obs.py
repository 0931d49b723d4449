import os, sys, ctypes
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pomcpp_b200 as pb
import oracle
orc = oracle.restatement()
n = 20011
mask, view = 1, 4
a = pb.Batch(n, n_templates=64, max_ticks=25)
c = pb.Batch(n, n_templates=64, max_ticks=25)
stride = int(pb.lib().pom_batch_obs_stride(a.h))
obs = a.alloc(stride * pb.OBS_BYTES)
mv = a.alloc(4 * n)
flags = pb.STEP_AUTORESET | pb.STEP_COUNT
for t in range(14):
    a.generate_moves(mv, 5, t, 6); a.step_observe(mv, obs, mask, view, flags); a.sync()
    c.generate_moves(mv, 5, t, 6); c.step(mv, flags); c.sync()
    fused = np.zeros((1, stride, pb.OBS_BYTES), np.uint8)
    pb.lib().pom_device_copy(0, fused.ctypes.data_as(ctypes.c_void_p), obs, fused.nbytes)
    plain = c.observe_planes(mask, view)
    S, _ = c.download()
    want = orc.observe_planes_batch(S, 0, view)
    bf = (fused[0, :n] != want); bp = (plain[0] != want)
    print("tick", t, "fused!=def envs", int(bf.any(1).sum()), "plain!=def envs", int(bp.any(1).sum()))
    if bf.any():
        e = int(np.nonzero(bf.any(1))[0][0]); print(" fused env", e, "bytes", np.nonzero(bf[e])[0][:20], fused[0, e][bf[e]][:20], want[e][bf[e]][:20])
    if bp.any():
        e = int(np.nonzero(bp.any(1))[0][0]); print(" plain env", e, "bytes", np.nonzero(bp[e])[0][:20], plain[0, e][bp[e]][:20], want[e][bp[e]][:20])
