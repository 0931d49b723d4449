# which box is this, and how do the two per-tick kernels run on it?  (boxes of the pool differ by about 2 %)
nvidia-smi --query-gpu=name,pci.bus_id,clocks.max.sm,clocks.max.mem,power.limit,temperature.gpu,ecc.mode.current --format=csv
python - <<'PY'
import torch
p = torch.cuda.get_device_properties(0)
print("SMs", p.multi_processor_count, "L2", p.L2_cache_size, "mem", p.total_memory)
PY
for k in ws tile; do echo "kernel=$k"; POM_STEP_KERNEL=$k python tools/prof_step.py | grep "200 ticks"; done
echo "generic images"; POM_WS_NOSPEC=1 python tools/prof_step.py | grep "200 ticks"
