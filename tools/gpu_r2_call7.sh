#!/bin/bash
# BASELINE.json configs[4], drop-in semantics: the 15991x11993 picture, of which the reference's limit keeps the top-left 8192x8192 (65,536 CTUs)
mkdir -p gpurun_out
python bench.py --config 5 --mode crop --single-pass > gpurun_out/r2g_bench_config5_crop.json 2> gpurun_out/r2g_bench_config5_crop.err; echo "config5 crop rc=$?"
tail -c 1500 gpurun_out/r2g_bench_config5_crop.json; tail -3 gpurun_out/r2g_bench_config5_crop.err
