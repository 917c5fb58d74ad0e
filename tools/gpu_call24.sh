#!/bin/bash
# GPU call 24 (last GPU seconds of the round): bench.py on the reference arm's sample (first 4 batches = 40 steps) -- checks the new
# details.per_batch_ms / details.on_reference_sample keys
mkdir -p gpurun_out
timeout 110 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/c24_bench_first40.json 2> gpurun_out/c24_bench_first40.err
echo "rc $?"
python - <<'PY'
import json
try:
    b=[json.loads(l) for l in open('gpurun_out/c24_bench_first40.json') if l.startswith('{')][-1]
    print('value', b['value'], 'steps', b['steps'], 'details', {k: b['details'].get(k) for k in ('per_batch_ms','on_reference_sample','continuation_steps_taken')}, 'frac', b['roofline']['frac'])
except Exception as e:
    print('no line', e)
PY
tail -3 gpurun_out/c24_bench_first40.err | cut -c1-300
