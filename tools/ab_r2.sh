#!/bin/bash
# round-2 A/B: each library variant on the dry and the wet Model204 workload (kernel-only, 1 M links)
for L in "$@"; do
  for W in 0.0 1.0; do
    HLM_B200_LIB=$PWD/tiger_hlm_gpu_b200/$L timeout 300 python bench.py --links-per-gpu ${LINKS:-1000000} --steps 3 --warmup 3 --no-baselines --no-e2e --wet-fraction $W ${ARGS:-} 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('%-24s wet %.1f steps/s %.4e  kernel_ms %.3f  frac %.4f  att/acc %.4f  status %s' % ('$L', $W, d['value'], r['kernel_ms_avg'], r['frac'], d['attempts_per_accepted'], d['link_status_after_run']))"
  done
done
