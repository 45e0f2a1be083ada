#!/bin/bash
# Round-2 FINAL evidence (final binary of the round): ncu --set full of the dominant kernel of every bench workload at the bench's own
# sizes (plus eval_kernel and the one-iteration launch behind rlrm_iterate). Each command first runs plain (must exit 0; its
# JSON line gives the active agent-steps per launch), then once under ncu. Run under gpurun (1 GPU).
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
  rm -f gpurun_out/${name}.ncu-rep   # the reports are ~10 MB each; gpurun brings back at most 64 MiB
  echo "$name done"
}
cap r02c_qrm4_cfg3         train_qrm4_kernel      1 $B --workload cfg3
cap r02c_qrm4_cfg5_tables  train_qrm4_kernel      1 $B --workload cfg5_tables
cap r02c_ql_fast_cfg3      train_ql_fast          1 $B --workload cfg3_ql
cap r02c_ql_fast_cfg2      train_ql_fast          1 $B --workload cfg2_batch
cap r02c_qrmn5_cfg2        train_qrmn_kernel      1 $B --workload cfg2_batch_qrm
cap r02c_qrm_block_exp6    train_qrm_block        1 $B --workload ow_exp6_qrm
cap r02c_qrm_block_chain12 train_qrm_block        1 $B --workload cfg4_qrm
cap r02c_qlambda_sparse    train_qlambda_sparse   1 $B --workload cfg4
cap r02c_qlambda_dense     train_qlambda_kernel   1 $B --workload cfg4_dense
cap r02c_shared_1m         shared_train_kernel    1 $B --workload cfg5_shared --iters 64
cap r02c_shared_131k       shared_train_kernel    1 $B --workload cfg5_shared --instances 131072 --iters 64
cap r02c_shared_ow12       shared_train_kernel    1 $B --workload ow12_shared
cap r02c_qrm_block_cfg3_f64 train_qrm_block     1 $B --workload cfg3_f64
cap r02c_generic_ql_cfg3_f64 "^train_kernel$"  1 $B --workload cfg3_ql_f64
cap r02c_eval_cfg3         eval_kernel            1 python profiles/scripts/run_eval.py
# shared_train_cluster_kernel (cooperative launch WITH a cluster dimension) is not captured: ncu's kernel replay relaunches it without
# the cluster attribute ("LaunchFailed", then an illegal address in map_shared_rank) — its numbers come from bench lines only
cap r02c_iterate_cfg3      "^train_kernel$"       100 python profiles/scripts/run_iterate.py
ls -la gpurun_out | head -60
