#!/bin/bash
# round-2 GPU pass F (1 GPU): integer-valued bound / multi_dot, full suite, bench, spans
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2f_t1.log 2>&1
echo "t1 rc=$?"; tail -4 gpurun_out/r2f_t1.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
echo "bench rc=$?"; tail -2 gpurun_out/r2f_bench.err; python - <<PY
import json
try:
    b=json.loads([l for l in open('gpurun_out/r2f_bench.json') if l.startswith('{')][-1])
    print({k:b[k] for k in ('value','ms_per_step')}, b['e2e'], b['throughput_batched'], b['roofline']['frac'])
    for r in b['configs']: print({k:r[k] for k in ('name','densify_ms','commit_ms','prove_ms','golden_match')})
except Exception as e: print('no bench line', e)
PY
for cfg in xor_c4_s20 lt_c8_s22 rc40_c4_s24; do timeout 600 python tools/spans_config.py $cfg 2 > gpurun_out/r2f_spans_$cfg.log 2>&1; tail -2 gpurun_out/r2f_spans_$cfg.log; done
