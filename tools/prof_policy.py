import sys; sys.path.insert(0,'/root/repo')
import pomcpp_b200 as pb
n=1<<18
b=pb.Batch(n,n_templates=4096,max_ticks=800)
b.rollout(150,99,0,pb.ROLL_SIMPLE(15))   # develop games (kernel 1: k_rollout policy)
moves=b.alloc(4*n)
for t in range(3):
    b.policy_moves(moves,99,150+t,15)
    b.step(moves,pb.STEP_AUTORESET)
b.sync()
