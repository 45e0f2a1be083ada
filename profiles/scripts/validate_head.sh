#!/bin/bash
# full validation of the current tree on one B200: GPU test suite, fuzz soak, smoke(), the default bench line and the reference arm
set -u
mkdir -p gpurun_out
T=${1:-r02_tv}
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/${T}_pytest.log 2>&1; tail -4 gpurun_out/${T}_pytest.log
( time RLRM_FUZZ_CASES=${FUZZ:-2048} timeout 900 python -m pytest tests/test_fuzz_gpu.py -x -q -m gpu ) > gpurun_out/${T}_fuzz.log 2>&1; tail -4 gpurun_out/${T}_fuzz.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -3 gpurun_out/${T}_smoke.log
( time timeout 600 python bench.py ) > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; tail -4 gpurun_out/${T}_bench_default.err
( time timeout 300 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; tail -4 gpurun_out/${T}_bench_reference.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/${T}_bench_default.json') if l.startswith('{')][-1])
print('headline', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'dram_frac', d['roofline']['dram_frac'])
for k,v in d['configs'].items(): print(k, v.get('value'), v.get('e2e_over_value'), v.get('error'))
print('call_by_call', d['e2e_call_by_call'] and d['e2e_call_by_call'].get('value'))
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline'].get('reference_python',{}).get('value'))
PY
