"""per-tick kernel time with and without POM_STEP_OVERLAP (device-resident moves, CUDA events)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pomcpp_b200 as pb
n = 1 << 20
b = pb.Batch(n, n_templates=4096, max_ticks=800)
ring = 64
moves = b.alloc(4 * n * ring)
for t in range(ring):
    b.generate_moves(moves.value + 4 * n * t, 11, 1000 + t, 6)
b.rollout(96, 7, 0, 0)
for name, extra in (("plain", 0), ("overlap", pb.STEP_OVERLAP), ("plain", 0), ("overlap", pb.STEP_OVERLAP)):
    fl = pb.STEP_AUTORESET | pb.STEP_COUNT | extra
    for t in range(20):
        b.step(moves.value + 4 * n * (t % ring), fl)
    b.sync(); b.event(0)
    K = 1000
    for t in range(K):
        b.step(moves.value + 4 * n * (t % ring), fl)
    b.event(1); b.sync()
    ms = b.elapsed_ms() / K
    print("%-8s %.4f ms/tick  %.3e env-steps/s  roofline %.3f" % (name, ms, n / ms * 1e3, 582 * n / (ms * 1e-3) / 6546.2e9))
