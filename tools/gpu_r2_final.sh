#!/bin/bash
# final validation of the head: smoke, full GPU suite, default bench (both arms)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2x_smoke.log 2>&1; tail -1 gpurun_out/r2x_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2x_gputests.log 2>&1; tail -2 gpurun_out/r2x_gputests.log
python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/r2x_bench_reference.json 2> gpurun_out/r2x_bench_reference.err; tail -c 300 gpurun_out/r2x_bench_reference.json
python bench.py --steps 5 --warmup 3 > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2x_bench.json").read().strip().splitlines()[-1])
print("value %.2f e2e %.2f (%.1f %%) frac %.4f cpu %.3f" % (d["value"], d["e2e"]["value"], 100 * d["e2e"]["value"] / d["value"], d["roofline"]["frac"], d["cpu_baseline"]["value"]))
PY
