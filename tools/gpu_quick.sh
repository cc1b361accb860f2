#!/bin/bash
# Quick GPU check: parity tests + short bench (device-resident only) + pipe peaks.
export PYTHONPATH=.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
timeout 100 ./tools/peaks > gpurun_out/peaks.json 2>&1; cat gpurun_out/peaks.json
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_quick.json'))
    print("ms_per_step %.2f value %.3e"%(d['ms_per_step'], d['value']))
    for k,v in d['kernels'].items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items()})
    print(d['extra'])
except Exception as e:
    print("bench parse failed", e); print(open('gpurun_out/bench_quick.err').read()[-3000:])
PY
