#!/bin/bash
# A/B a set of library builds on the GPU box: tools/ab.sh lib1.so lib2.so ...   (paths under tiger_hlm_gpu_b200/)
for L in "$@"; do
  HLM_B200_LIB=$PWD/tiger_hlm_gpu_b200/$L timeout 300 python bench.py --links-per-gpu ${LINKS:-1000000} --steps 3 --warmup 2 --no-baselines --no-e2e ${EXTRA} 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('%-28s steps/s %.4e  kernel_ms %.3f  attempts/s %.4e  att/acc %.4f  status %s' % ('$L', d['value'], r['kernel_ms_avg'], r['attempts_per_launch']/r['kernel_ms_avg']*1e3, d['attempts_per_accepted'], d['link_status_after_run']))"
done
