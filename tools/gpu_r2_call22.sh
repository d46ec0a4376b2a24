#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_boundary.py tests/test_gpu_configs.py -q -x > gpurun_out/r2u_tests.log 2>&1; tail -3 gpurun_out/r2u_tests.log
python bench.py --config 3 --total 8192 --steps 2 --warmup 3 --no-cpu > gpurun_out/r2u_bench_config3.json 2> gpurun_out/r2u_bench_config3.err; echo "config3 rc=$?"; tail -2 gpurun_out/r2u_bench_config3.err
python bench.py --steps 3 --warmup 3 > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("bench_config3", "bench"):
    d = json.loads(open(f"gpurun_out/r2u_{f}.json").read().strip().splitlines()[-1])
    print(f, "value %.2f e2e %.2f (%.1f %%)" % (d["value"], d["e2e"]["value"], 100 * d["e2e"]["value"] / d["value"]))
PY
