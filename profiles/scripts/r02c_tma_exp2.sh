# TMA fetch variant in the dense-batch regime: twice / four times the instances of the bench defaults
mkdir -p gpurun_out
for v in base tma; do
for spec in "cfg3_f64 131072" "cfg3_f64 262144" "ow_exp6_qrm 131072" "ow_exp6_qrm 262144" "cfg4_qrm 32768" "cfg4_qrm 131072"; do
set -- $spec
RLRM_LIB_PATH=$PWD/build/variants/librlrm_$v.so timeout 200 python bench.py --no-cpu-baseline --no-configs --no-call-by-call --steps 5 --warmup 3 --iters 512 --workload $1 --instances $2 2>/dev/null | \
  python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v $1 $2', '%.4e' % d['value'], '%.3f ms' % d['ms_per_step'])"
done; done | tee gpurun_out/r02_t40_tma_dense.txt
