# sparse Q(lambda): lanes per agent (RLRM_QLS_LG) A/B on BASELINE config 4 + parity under the alternative layout
mkdir -p gpurun_out
for lg in 4; do
RLRM_QLS_LG=$lg timeout 900 python -m pytest tests -x -q -m gpu -k "qlambda or sparse or fuzz or golden or checkpoint" 2>&1 | tail -2
done
for lg in 32 4; do
for st in 10 50; do
RLRM_QLS_LG=$lg timeout 300 python bench.py --workload cfg4 --steps $st --warmup 3 --no-cpu-baseline --no-configs --no-call-by-call 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('lg=$lg steps=$st', '%.4e'%d['value'], '%.3f ms'%d['ms_per_step'], 'L=%.1f'%d['roofline']['mean_live_traces'], 'e2e %.3e'%d['e2e']['value'])"
done; done
