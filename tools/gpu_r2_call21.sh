#!/bin/bash
# evidence for the head commit: launch list + full ncu capture of the throughput variant, a lighter capture of the cluster variant,
# configs[2] as written (8192 pictures) with the final kernel
mkdir -p gpurun_out
python tools/prof_run.py 1036 96 128 2 2 > gpurun_out/r2t_plain_g7.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 20 --csv --log-file gpurun_out/r2t_launches.csv python tools/prof_run.py 1036 96 128 2 2 > gpurun_out/r2t_ncu1.log 2>&1
timeout 600 ncu --set full --metrics smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,smsp__inst_executed_op_shared_ld.sum,smsp__inst_executed_op_shared_st.sum,smsp__inst_executed_op_global_ld.sum,smsp__inst_executed_op_global_st.sum --clock-control none --import-source on -k regex:hevce_encode -s 1 -c 1 -o gpurun_out/r2t_prof_g7 python tools/prof_run.py 1036 96 128 2 2 > gpurun_out/r2t_ncu2.log 2>&1
cat gpurun_out/r2t_plain_g7.log
python tools/prof_run.py 74 96 128 2 2 > gpurun_out/r2t_plain_c2.log 2>&1 &&
timeout 300 ncu --set full --clock-control none -k regex:hevce_encode -s 1 -c 1 -o gpurun_out/r2t_prof_c2 python tools/prof_run.py 74 96 128 2 2 > gpurun_out/r2t_ncu3.log 2>&1
cat gpurun_out/r2t_plain_c2.log; tail -2 gpurun_out/r2t_ncu3.log
python bench.py --config 3 --total 8192 --steps 2 --warmup 3 --no-cpu > gpurun_out/r2t_bench_config3.json 2> gpurun_out/r2t_bench_config3.err; echo "config3 rc=$?"; tail -2 gpurun_out/r2t_bench_config3.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2t_bench_config3.json").read().strip().splitlines()[-1])
print("config3 8192: value %.2f e2e %.2f (%.1f %%)" % (d["value"], d["e2e"]["value"], 100 * d["e2e"]["value"] / d["value"]))
PY
