#!/bin/bash
# parent || child variant t1 on the GPU: parity per variant, latency per variant, racecheck of a 7-picture gang and of the track variant
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_variants.py tests/test_gpu_boundary.py -q -s > gpurun_out/r2d_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_gputests.log
tail -6 gpurun_out/r2d_gputests.log
python tools/variant_bench.py 64 64 2 > gpurun_out/r2d_variants_q2.log 2>&1; cat gpurun_out/r2d_variants_q2.log
python tools/variant_bench.py 64 64 4 w1 t1 > gpurun_out/r2d_variants_q4.log 2>&1; cat gpurun_out/r2d_variants_q4.log
python tools/variant_bench.py 256 256 2 w1 t1 > gpurun_out/r2d_variants_256.log 2>&1; cat gpurun_out/r2d_variants_256.log
HEVCE_VARIANT=g7 timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/prof_run.py 7 64 64 2 1 > gpurun_out/r2d_racecheck_g7.log 2>&1; tail -5 gpurun_out/r2d_racecheck_g7.log
HEVCE_VARIANT=t1 timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/prof_run.py 1 64 64 2 1 > gpurun_out/r2d_racecheck_t1.log 2>&1; tail -5 gpurun_out/r2d_racecheck_t1.log
