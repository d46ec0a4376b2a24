#!/bin/bash
mkdir -p gpurun_out
python tools/ab_variants.py run g7 64 64 2 2>&1 | tee gpurun_out/r2n_ab_fastrel.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2n_gputests.log 2>&1; tail -3 gpurun_out/r2n_gputests.log
python tools/variant_bench.py 64 64 2 > gpurun_out/r2n_variants.log 2>&1; cat gpurun_out/r2n_variants.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2n_bench.json').read().strip().splitlines()[-1])
print('value %.2f e2e %.2f (%.1f %%) ms/step %.0f frac %.4f' % (d['value'], d['e2e']['value'], 100*d['e2e']['value']/d['value'], d['ms_per_step'], d['roofline']['frac']))
PY
tail -2 gpurun_out/r2n_bench.err
