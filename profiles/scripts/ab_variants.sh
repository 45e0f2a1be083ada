#!/bin/bash
# A/B of library builds: every build/variants/*.so runs the listed bench workloads (device-timed value only)
set -u
mkdir -p gpurun_out
WL=${WL:-"cfg3 cfg3_ql cfg2_batch cfg5_tables ow_exp6_qrm cfg3_f64"}
REP=${REP:-2}
for rep in $(seq 1 $REP); do
for v in build/variants/*.so; do
  for w in $WL; do
    RLRM_LIB_PATH=$PWD/$v timeout 200 python bench.py --no-cpu-baseline --no-configs --no-call-by-call --steps 8 --warmup 3 --workload $w 2>/dev/null | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$(basename $v) $w rep$rep', '%.4e' % d['value'], '%.3f ms' % d['ms_per_step'])"
  done
done
done | tee gpurun_out/${TAG:-ab}_variants.txt
