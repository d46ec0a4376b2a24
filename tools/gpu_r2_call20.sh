#!/bin/bash
# full GPU suite with all six variants, then the large configs with the cluster variant
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2s_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s_gputests.log; tail -4 gpurun_out/r2s_gputests.log
for q in 4 0; do
timeout 600 python bench.py --config 4 --qpd6 $q --single-pass > gpurun_out/r2s_bench_config4_q$q.json 2> gpurun_out/r2s_bench_config4_q$q.err; echo "config4 q$q rc=$?"; tail -2 gpurun_out/r2s_bench_config4_q$q.err
done
timeout 900 python bench.py --config 5 --mode crop --single-pass > gpurun_out/r2s_bench_config5_crop.json 2> gpurun_out/r2s_bench_config5_crop.err; echo "config5 crop rc=$?"
python - <<'PY'
import json
for f in ("config4_q4", "config4_q0", "config5_crop"):
    try:
        d = json.loads(open(f"gpurun_out/r2s_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, "value %.3f e2e %.3f ms %.0f variant %s grid %s gold %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["config"]["kernel_variant"], d["config"]["grid_ctas"], d["parity"]["checked_against_reference_manifest"]))
    except Exception as e:
        print(f, "FAILED", e)
PY
