#!/bin/bash
# BASELINE.json configs[4], raised limit: 15991x11993 -> 16000x12000 (187,500 CTUs) on one CTA, against the reference built with its two limits raised
mkdir -p gpurun_out
python bench.py --config 5 --mode xl --single-pass > gpurun_out/r2q_bench_config5_xl.json 2> gpurun_out/r2q_bench_config5_xl.err; echo "config5 xl rc=$?"
tail -c 900 gpurun_out/r2q_bench_config5_xl.json; tail -3 gpurun_out/r2q_bench_config5_xl.err
