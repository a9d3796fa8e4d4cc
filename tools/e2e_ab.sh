for cfg in "zc" "st4" "st8" "st16"; do
  case $cfg in zc) export DEXR_HOST_ZEROCOPY=1; unset DEXR_HOST_CHUNKS;; st4) export DEXR_HOST_ZEROCOPY=0 DEXR_HOST_CHUNKS=4;; st8) export DEXR_HOST_ZEROCOPY=0 DEXR_HOST_CHUNKS=8;; st16) export DEXR_HOST_ZEROCOPY=0 DEXR_HOST_CHUNKS=16;; esac
  python bench.py --steps 10 --warmup 3 --no-configs --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print('$cfg', 'e2e %.4e (%.3f ms)  ref_value form %.4e (%.3f ms)  pageable %.3e'%(e['value'],e['ms_per_step'],e['ref_value_form']['value'],e['ref_value_form']['ms_per_step'],e['staged_pageable']['value']))"
done
