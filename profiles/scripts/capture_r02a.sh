#!/bin/bash
# Round-2 "before" evidence: ncu --set full of the kernels round 1 left uncaptured, on the
# round-1 binary. Each bench command first runs plain (must exit 0), then once under ncu. Run under gpurun (1 GPU).
set -u
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --steps 2 --warmup 1"
cap() {  # name, kernel regex, bench args
  local name=$1 kern=$2; shift 2
  $B "$@" > gpurun_out/${name}_plain.json 2> gpurun_out/${name}_plain.err || { echo "$name plain run failed"; return; }
  ncu --set full --clock-control none --import-source on -k "regex:$kern" -s 1 -c 1 -f -o gpurun_out/${name} \
      $B "$@" > gpurun_out/${name}_ncu.log 2>&1
  ncu -i gpurun_out/${name}.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  ncu -i gpurun_out/${name}.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/${name}_source.csv.gz
  rm -f gpurun_out/${name}.ncu-rep   # the reports are ~10 MB each; gpurun brings back at most 64 MiB
}
cap r02a_qrm4_cfg3       train_qrm4_kernel   --workload cfg3 --iters 512
cap r02a_qrm4_cfg5size   train_qrm4_kernel   --workload cfg5_tables --instances 262144 --iters 128
cap r02a_ql_fast_cfg3    train_ql_fast       --workload cfg3_ql --iters 512
cap r02a_generic_exp6    train_kernel        --workload ow_exp6_qrm --iters 256
cap r02a_generic_chain12 train_kernel        --workload cfg4_qrm --iters 128
cap r02a_qlambda_dense   train_qlambda_kernel --workload cfg4_dense --instances 16384 --iters 4
cap r02a_shared_propose  shared_propose      --workload cfg5_shared --instances 262144 --iters 16
ls -la gpurun_out
