"""where does the end-to-end loop lose time against the device-resident tick?  The bench's e2e loop (two half-batches stepped
alternately through pom_batch_step_compact, the host waits for every half-batch every tick) with the buffers in pinned
host memory or in device memory, with and without outputs, and without the per-tick waits."""
import os, sys, time
import ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pomcpp_b200 as pb

n = 1 << 20
h = n // 2
K = int(os.environ.get("POM_PROBE_STEPS", 400))
RING = 32
fl = pb.STEP_AUTORESET | pb.STEP_COUNT
H = [pb.Batch(h, env_offset=i * h, n_templates=4096, max_ticks=800) for i in range(2)]
for x in H:
    x.rollout(96, 5, 0, 0)
    x.sync()
rng = np.random.default_rng(1)
keep = []
def pinned(shape, dt):
    a, o = pb.pinned_array(shape, dt, near_device=0)
    keep.append(o)
    return a
joint_h = [[pinned((h,), np.uint16) for _ in range(RING)] for _ in range(2)]
for i in range(2):
    for a in joint_h[i]:
        a[:] = rng.integers(0, 1296, size=h, dtype=np.uint16)
joint_d = [[None] * RING for _ in range(2)]
for i in range(2):
    for k in range(RING):
        p = H[i].alloc(2 * h)
        pb._ck(pb.lib().pom_device_copy(0, p, joint_h[i][k].ctypes.data_as(C.c_void_p), 2 * h))
        joint_d[i][k] = p.value
bits_h = [pinned(((h + 31) // 32,), np.uint32) for _ in range(2)]
fenv_h = [pinned((h,), np.uint32) for _ in range(2)]
fst_h = [pinned((h,), np.uint8) for _ in range(2)]
fcnt_h = [pinned((1,), np.uint32) for _ in range(2)]
bits_d = [H[i].alloc(4 * ((h + 31) // 32)).value for i in range(2)]
fenv_d = [H[i].alloc(4 * h).value for i in range(2)]
fst_d = [H[i].alloc(h).value for i in range(2)]
fcnt_d = [H[i].alloc(4).value for i in range(2)]

stride = int(pb.lib().pom_batch_obs_stride(H[0].h))
obs_d = [H[i].alloc(stride * pb.OBS_BYTES) for i in range(2)]

def loop(name, jin, outs, wait=True, steps=K, obs=False):
    ios = [[H[i].compact_io(jin[i][k], *(outs[i] if outs else (None, None, None, None)), obs_d[i] if obs else None, 1 if obs else 0, 4)
            for k in range(RING)] for i in range(2)]
    if outs and outs[0][1] is not None:
        for i in range(2):
            for io in ios[i]:
                io.fin_capacity = h
    def go(i, k):
        H[i].step_compact_io(ios[i][k % RING], fl)
    def run(steps):
        go(0, 0)
        for k in range(steps):
            go(1, k)
            if wait: H[0].sync()
            if k + 1 < steps: go(0, k + 1)
            if wait: H[1].sync()
        H[0].sync(); H[1].sync()
    run(10)
    t0 = time.perf_counter()
    run(steps)
    dt = (time.perf_counter() - t0) / steps
    print("%-64s %.1f us per tick  %.3e env-steps/s" % (name, dt * 1e6, n / dt), flush=True)

outs_h = [(bits_h[i], fenv_h[i], fst_h[i], fcnt_h[i]) for i in range(2)]
outs_d = [(bits_d[i], fenv_d[i], fst_d[i], fcnt_d[i]) for i in range(2)]
loop("pinned joint, pinned outputs, wait per half (= bench e2e)", joint_h, outs_h)
loop("pinned joint, pinned done bits only, wait per half", joint_h, [(bits_h[i], None, None, None) for i in range(2)])
loop("pinned joint, pinned finished-env list only, wait per half", joint_h, [(None, fenv_h[i], fst_h[i], fcnt_h[i]) for i in range(2)])
loop("pinned joint, device done bits + pinned list, wait per half", joint_h, [(bits_d[i], fenv_h[i], fst_h[i], fcnt_h[i]) for i in range(2)])
loop("device joint, pinned outputs, wait per half", joint_d, outs_h)
loop("pinned joint, no outputs, wait per half", joint_h, None)
loop("device joint, device outputs, wait per half", joint_d, outs_d)
loop("device joint, no outputs, wait per half", joint_d, None)
loop("pinned joint, pinned outputs, no waits (queue runs ahead)", joint_h, outs_h, wait=False)
loop("device joint, device outputs, no waits", joint_d, outs_d, wait=False)
loop("device joint, no outputs, no waits", joint_d, None, wait=False)
loop("OBS: pinned joint, pinned outputs, wait per half (= bench e2e_obs)", joint_h, outs_h, obs=True, steps=K // 2)
loop("OBS: device joint, no outputs, wait per half", joint_d, None, obs=True, steps=K // 2)
loop("OBS: device joint, no outputs, no waits", joint_d, None, wait=False, obs=True, steps=K // 2)
loop("OBS: pinned joint, pinned outputs, no waits", joint_h, outs_h, wait=False, obs=True, steps=K // 2)
# one handle alone, compact path, with and without observation planes (is it the alternation or the path?)
for obs in (False, True):
    ios = [H[0].compact_io(joint_d[0][k], None, None, None, None, obs_d[0] if obs else None, 1 if obs else 0, 4) for k in range(RING)]
    for k in range(10):
        H[0].step_compact_io(ios[k % RING], fl)
    H[0].sync()
    H[0].event(0)
    for k in range(100):
        H[0].step_compact_io(ios[k % RING], fl)
    H[0].event(1)
    H[0].sync()
    print("one half alone, compact, obs=%s: %.1f us per launch" % (obs, H[0].elapsed_ms() * 10), flush=True)
# the same actions as 4 move bytes per env
mv4_d = []
for k in range(RING):
    j = joint_h[0][k].astype(np.uint32)
    m4 = np.stack([j % 6, (j // 6) % 6, (j // 36) % 6, (j // 216) % 6], axis=1).astype(np.uint8)
    p4 = H[0].alloc(4 * h)
    pb._ck(pb.lib().pom_device_copy(0, p4, np.ascontiguousarray(m4).ctypes.data_as(C.c_void_p), 4 * h))
    mv4_d.append(p4)
for obs in (False, True):
    f = (lambda k: H[0].step_observe(mv4_d[k % RING], obs_d[0], 1, 4, fl)) if obs else (lambda k: H[0].step(mv4_d[k % RING], fl))
    for k in range(10):
        f(k)
    H[0].sync()
    H[0].event(0)
    for k in range(100):
        f(k)
    H[0].event(1)
    H[0].sync()
    print("one half alone, 4-byte moves (same actions), obs=%s: %.1f us per launch" % (obs, H[0].elapsed_ms() * 10), flush=True)
