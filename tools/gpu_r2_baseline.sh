#!/bin/bash
# round-2 baseline of the head kernel on the bench workload: tests, bench, launch list, one full ncu capture
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_gputests.log
tail -3 gpurun_out/r2a_gputests.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2a_bench.json
python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/r2a_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/r2a_ncu1.log 2>&1
python tools/prof_run.py 1036 512 768 2 2 > gpurun_out/r2a_plain2.log 2>&1 &&
ncu --set full --metrics smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,smsp__inst_executed_op_shared_ld.sum,smsp__inst_executed_op_shared_st.sum,smsp__inst_executed_op_global_ld.sum,smsp__inst_executed_op_global_st.sum --clock-control none --import-source on -k regex:hevce_encode -s 1 -c 1 -o gpurun_out/r2a_prof python tools/prof_run.py 1036 512 768 2 2 > gpurun_out/r2a_ncu2.log 2>&1
tail -3 gpurun_out/r2a_ncu2.log
ls -la gpurun_out
