#!/bin/bash
# full GPU suite on the multi-variant library, default bench, configs[2] as written (8192 pictures), configs[3] at both QPs,
# and the ncu evidence on a workload ncu can replay (1036 pictures of 128x96: every CTA holds a full gang of 7)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2c_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_gputests.log
tail -4 gpurun_out/r2c_gputests.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"; tail -c 700 gpurun_out/r2c_bench.json; tail -2 gpurun_out/r2c_bench.err
python bench.py --config 3 --total 8192 --steps 2 --warmup 3 --no-cpu > gpurun_out/r2c_bench_config3.json 2> gpurun_out/r2c_bench_config3.err; echo "config3 rc=$?"; tail -c 900 gpurun_out/r2c_bench_config3.json; tail -2 gpurun_out/r2c_bench_config3.err
for q in 4 0; do
python bench.py --config 4 --qpd6 $q --single-pass > gpurun_out/r2c_bench_config4_q$q.json 2> gpurun_out/r2c_bench_config4_q$q.err; echo "config4 q$q rc=$?"; tail -c 900 gpurun_out/r2c_bench_config4_q$q.json; tail -2 gpurun_out/r2c_bench_config4_q$q.err
done
python tools/prof_run.py 1036 96 128 2 2 > gpurun_out/r2c_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 20 --csv --log-file gpurun_out/r2c_launches.csv python tools/prof_run.py 1036 96 128 2 2 > gpurun_out/r2c_ncu1.log 2>&1
timeout 600 ncu --set full --metrics smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,smsp__inst_executed_op_shared_ld.sum,smsp__inst_executed_op_shared_st.sum,smsp__inst_executed_op_global_ld.sum,smsp__inst_executed_op_global_st.sum --clock-control none --import-source on -k regex:hevce_encode -s 1 -c 1 -o gpurun_out/r2c_prof python tools/prof_run.py 1036 96 128 2 2 > gpurun_out/r2c_ncu2.log 2>&1
tail -3 gpurun_out/r2c_ncu2.log; cat gpurun_out/r2c_plain.log
