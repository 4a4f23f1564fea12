#!/bin/bash
# A/B of library variants on the two lane-schedule workloads: the routed hour (2.5 M links) and Model 200 unrouted (1 M links)
for L in "$@"; do
  HLM_B200_LIB=$PWD/tiger_hlm_gpu_b200/$L timeout 600 python bench.py --workload routed --steps 5 --warmup 3 --no-e2e ${ARGS:-} 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('%-24s routed   steps/s %.4e  ms/step %.3f  kernel_ms %.3f  frac %.4f  att/acc %.4f' % ('$L', d['value'], d['ms_per_step'], r['kernel_ms_avg'], r['frac'], d['attempts_per_accepted']))"
  HLM_B200_LIB=$PWD/tiger_hlm_gpu_b200/$L timeout 600 python bench.py --workload model200 --steps 5 --warmup 3 --no-e2e --no-baselines ${ARGS:-} 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('%-24s model200 steps/s %.4e  ms/step %.3f  kernel_ms %.3f  frac %.4f  att/acc %.4f' % ('$L', d['value'], d['ms_per_step'], r['kernel_ms_avg'], r['frac'], d['attempts_per_accepted']))"
done
