set -x
for n in 131072 1048576; do
for cl in 0 2 4 8; do
RLRM_SHARED_CLUSTER=$cl timeout 300 python bench.py --workload cfg5_shared --instances $n --iters 256 --steps 5 --warmup 3 --no-cpu-baseline --no-configs --no-call-by-call > gpurun_out/r02_t21_shared_cl${cl}_${n}.json 2> gpurun_out/r02_t21_shared_cl${cl}_${n}.err
python -c "
import json;d=json.load(open('gpurun_out/r02_t21_shared_cl${cl}_${n}.json'));print('cl',$cl,'n',$n,d['value'],d['ms_per_step'],d['roofline'].get('kernel'))"
done; done
