# float64 tables on BASELINE config 3: specialised block kernel vs the generic kernel (RLRM_FORCE_GENERIC=1), QRM and plain QL
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "float64 or f64 or generic_kernel_equals" > gpurun_out/r02_t23_pytest.log 2>&1; tail -3 gpurun_out/r02_t23_pytest.log
for g in 0 1; do for w in cfg3_f64 cfg3_ql_f64; do
if [ $g = 1 ]; then export RLRM_FORCE_GENERIC=1; else unset RLRM_FORCE_GENERIC; fi
timeout 180 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-configs --no-call-by-call > gpurun_out/r02_t23_${w}_g$g.json 2> gpurun_out/r02_t23_${w}_g$g.err; tail -2 gpurun_out/r02_t23_${w}_g$g.err
python -c "
import json;d=json.load(open('gpurun_out/r02_t23_${w}_g$g.json'));print('$w generic=$g',d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d['dtype'])"
done; done
