#!/bin/bash
# Round-2 (late) evidence: ncu --set full of train_qrm_block_kernel<FrozenLake,3,double> (BASELINE config 3 on the reference's own
# float64 tables) at the bench's size; same recipe as capture_r02b.sh. Optional A/B of the register budget (QRMB_MINB_F64).
set -u
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-configs --no-call-by-call --steps 2 --warmup 1"
cap() {  # name, kernel regex, launches to skip, command...
  local name=$1 kern=$2 skip=$3; shift 3
  "$@" > gpurun_out/${name}_plain.json 2> gpurun_out/${name}_plain.err || { echo "$name plain run failed"; tail -3 gpurun_out/${name}_plain.err; return; }
  ncu --set full --clock-control none --import-source on -k "regex:$kern" -s $skip -c 1 -f -o gpurun_out/${name} \
      "$@" > gpurun_out/${name}_ncu.log 2>&1
  ncu -i gpurun_out/${name}.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  ncu -i gpurun_out/${name}.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/${name}_source.csv.gz
  rm -f gpurun_out/${name}.ncu-rep
  echo "$name done"
}
for v in build/variants/*.so; do
  [ -f "$v" ] || continue
  RLRM_LIB_PATH=$PWD/$v python bench.py --no-cpu-baseline --no-configs --no-call-by-call --steps 5 --warmup 3 --workload cfg3_f64 2>/dev/null | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['value'], d['ms_per_step'])"
done
cap r02c_qrm_block_cfg3_f64 train_qrm_block 1 $B --workload cfg3_f64
cap r02c_generic_ql_cfg3_f64 "^train_kernel$" 1 $B --workload cfg3_ql_f64
