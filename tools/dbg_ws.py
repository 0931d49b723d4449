"""debug: persistent k_step_ws against the tile kernel k_step on identical inputs"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pomcpp_b200 as pb

def run(kernel, n, ticks, preroll, bulk=True):
    os.environ["POM_STEP_KERNEL"] = kernel
    b = pb.Batch(n, n_templates=4096, max_ticks=800)
    if preroll:
        b.rollout(preroll, 5, 0, 0)
    mv = b.alloc(4 * n + 64)
    base = mv.value if bulk else mv.value + 4
    for t in range(ticks):
        b.generate_moves(base, 77, t, 6)
        b.sync()
        try:
            b.step(base, pb.STEP_AUTORESET | pb.STEP_COUNT)
            b.sync()
        except Exception as e:
            raise RuntimeError("step at tick %d: %s" % (t, e))
    st = b.stats().as_dict()
    recs = b.records() if hasattr(b, "records") else None
    S, status = b.download(0, min(n, 200000))
    b.free(mv); b.close()
    return st, S, status

for n, ticks, pre in ((113664, 20, 0), (113664 + 4736, 20, 0), (1 << 18, 20, 0), (1 << 20, 8, 40), (1 << 20, 24, 0)):
    for bulk in (True, False):
        try:
            a = run("ws", n, ticks, pre, bulk)
        except Exception as e:
            print("n=%d ticks=%d pre=%d bulk=%s ws FAILED: %s" % (n, ticks, pre, bulk, e)); continue
        c = run("tile", n, ticks, pre, bulk)
        same = a[0] == c[0] and a[1].tobytes() == c[1].tobytes() and (a[2] == c[2]).all()
        print("n=%d ticks=%d pre=%d bulk=%s  ws==tile: %s  %s" % (n, ticks, pre, bulk, same, a[0] if not same else ""))
