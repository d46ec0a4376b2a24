#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_boundary.py -q -x > gpurun_out/r2i_tests.log 2>&1; tail -3 gpurun_out/r2i_tests.log
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2i_bench.json').read().strip().splitlines()[-1])
print('value %.2f e2e %.2f (%.1f %%) ms/step %.0f' % (d['value'], d['e2e']['value'], 100*d['e2e']['value']/d['value'], d['ms_per_step']))
PY
tail -2 gpurun_out/r2i_bench.err
python bench.py --images 2072 --steps 2 --warmup 3 --no-cpu > gpurun_out/r2i_bench2072.json 2> gpurun_out/r2i_bench2072.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2i_bench2072.json').read().strip().splitlines()[-1])
print('2072: value %.2f e2e %.2f (%.1f %%) ms/step %.0f' % (d['value'], d['e2e']['value'], 100*d['e2e']['value']/d['value'], d['ms_per_step']))
PY
