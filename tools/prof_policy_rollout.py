"""profiling helper: the fused rollout with four SimpleAgents (mid-game)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pomcpp_b200 as pb
b = pb.Batch(1 << 18, n_templates=4096, max_ticks=800)
b.rollout(150, 99, 0, pb.ROLL_SIMPLE(15))
b.sync()
b.rollout(40, 99, 150, pb.ROLL_SIMPLE(15))
b.sync()
