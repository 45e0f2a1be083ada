#!/bin/bash
# re-capture of the two generic-kernel launches after the agent_step closure fix (their r02c captures had the 656-byte local frame)
set -u
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-configs --no-call-by-call --steps 2 --warmup 1"
cap() {
  local name=$1 kern=$2 skip=$3; shift 3
  "$@" > gpurun_out/${name}_plain.json 2> gpurun_out/${name}_plain.err || { echo "$name plain run failed"; tail -3 gpurun_out/${name}_plain.err; return; }
  ncu --set full --clock-control none --import-source on -k "regex:$kern" -s $skip -c 1 -f -o gpurun_out/${name} \
      "$@" > gpurun_out/${name}_ncu.log 2>&1
  ncu -i gpurun_out/${name}.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  ncu -i gpurun_out/${name}.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/${name}_source.csv.gz
  rm -f gpurun_out/${name}.ncu-rep
  echo "$name done"
}
cap r02c_qrm_block_chain12 train_qrm_block 1 $B --workload cfg4_qrm

