#!/bin/bash
# final validation of the head after the coder / RDOQ diet: smoke, full GPU suite, default bench, launch list and a full ncu
# capture of the shipped throughput kernel (same workload as profiles/r2_ncu_summary.txt: 1036 pictures of 128x96, qpd6=2)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; tail -1 gpurun_out/r2z_smoke.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2z_gputests.log 2>&1; tail -2 gpurun_out/r2z_gputests.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2z_bench.json").read().strip().splitlines()[-1])
print("value %.2f e2e %.2f (%.1f %%) frac %.4f cpu %.3f" % (d["value"], d["e2e"]["value"], 100 * d["e2e"]["value"] / d["value"], d["roofline"]["frac"], d["cpu_baseline"]["value"]))
PY
python tools/prof_run.py 1036 96 128 2 2 > gpurun_out/r2z_plain_g7.log 2>&1 &&
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 20 --csv --log-file gpurun_out/r2z_launches.csv python tools/prof_run.py 1036 96 128 2 2 > gpurun_out/r2z_ncu1.log 2>&1
timeout 400 ncu --set full --metrics smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,smsp__inst_executed_op_shared_ld.sum,smsp__inst_executed_op_shared_st.sum,smsp__inst_executed_op_global_ld.sum,smsp__inst_executed_op_global_st.sum --clock-control none --import-source on -k regex:hevce_encode -s 1 -c 1 -o gpurun_out/r2z_prof_g7 python tools/prof_run.py 1036 96 128 2 2 > gpurun_out/r2z_ncu2.log 2>&1
cat gpurun_out/r2z_plain_g7.log; tail -2 gpurun_out/r2z_ncu2.log
