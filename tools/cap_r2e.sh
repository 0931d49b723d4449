set -x
mkdir -p gpurun_out
python tools/prof_rollout.py > gpurun_out/prof_rollout.log 2>&1 || exit 1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_rollout -s 1 -c 1 -f -o gpurun_out/prof_krollout_r2e python tools/prof_rollout.py > gpurun_out/ncu_rollout.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_expand -s 3 -c 1 -f -o gpurun_out/prof_kexpand_r2e python tools/prof_expand.py > gpurun_out/ncu_expand.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_step_ws -s 30 -c 2 -f -o gpurun_out/prof_kstep_r2e python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/ncu_step.log 2>&1
ls -la gpurun_out
