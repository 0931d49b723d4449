"""one-off, longer than the committed tests: the CUDA path against the compiled, unmodified reference, every field of every
env after every tick (tests/test_gpu_round2.py::_gpu_vs_compiled_reference), random and all-kick stress regimes"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle
import pomcpp_b200 as pb
from test_gpu_round2 import _gpu_vs_compiled_reference
oracle.build()
orc, ref = oracle.restatement(), oracle.reference()
tot = 0
for name, args in (("harmless (config 2)", (65536, 800, 5, 0, 4242)), ("random", (65536, 300, 6, 0, 777)), ("stress", (32768, 300, 6, 1, 2025))):
    t0 = time.time()
    compared, excluded = _gpu_vs_compiled_reference(pb, orc, ref, *args)
    tot += compared
    print("%-20s %d envs x %d ticks: %d env-steps compared field by field, %d excluded (reference undefined), 0 mismatches, %.0f s"
          % (name, args[0], args[1], compared, excluded, time.time() - t0), flush=True)
print("total", tot)
