#!/bin/bash
# k_rank with a persisting L2 window over the first X MB of the posting rows (USB_RANK_L2_MB).
mkdir -p gpurun_out
for mb in 0 32 64 96 128; do
  USB_RANK_L2_MB=$mb python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-legs > gpurun_out/bench_l2_$mb.json 2> gpurun_out/bench_l2_$mb.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_l2_$mb.json").read().strip().splitlines()[-1])
print("L2 window $mb MB:", {k: round(v, 1) for k, v in d["kernels_ms_per_step"].items() if k.startswith("k_")})
PY
done
python -c "
import torch
p = torch.cuda.get_device_properties(0)
print('L2', p.L2_cache_size)
"
