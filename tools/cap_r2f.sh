# final captures of round 2 (run under gpurun; each program has exited 0 without ncu first)
set -x
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_pre_ncu.log 2>&1 || exit 1
python tools/prof_rollout.py > gpurun_out/prof_rollout.log 2>&1 || exit 1
python tools/prof_expand.py > gpurun_out/prof_expand.log 2>&1 || exit 1
python tools/prof_obs.py > gpurun_out/prof_obs.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r2f.csv python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_step_ws -s 30 -c 2 -f -o gpurun_out/prof_kstep_r2f python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/ncu_step.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_rollout -s 1 -c 1 -f -o gpurun_out/prof_krollout_r2f python tools/prof_rollout.py > gpurun_out/ncu_rollout.log 2>&1
if [ "$1" = all ]; then
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_expand -s 3 -c 1 -f -o gpurun_out/prof_kexpand_r2f python tools/prof_expand.py > gpurun_out/ncu_expand.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_observe_planes --launch-skip 10 --launch-count 1 -f -o gpurun_out/prof_kobs_r2f python tools/prof_obs.py > gpurun_out/ncu_obs.log 2>&1
fi
ls -la gpurun_out | tail -20
