#!/bin/bash
# multi-GPU: the library's own split (one call, all devices) and the one-process-per-GPU bench
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_boundary.py -q -s -k "several_devices or left_alone" > gpurun_out/r2m_multidev_tests.log 2>&1; tail -3 gpurun_out/r2m_multidev_tests.log
python bench.py --one-call $N --steps 2 --warmup 2 > gpurun_out/r2m_bench_onecall_$N.json 2> gpurun_out/r2m_bench_onecall_$N.err; echo "one-call rc=$?"; tail -c 800 gpurun_out/r2m_bench_onecall_$N.json; tail -2 gpurun_out/r2m_bench_onecall_$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 3 --no-cpu > gpurun_out/r2m_bench_ranks_$N.json 2> gpurun_out/r2m_bench_ranks_$N.err; echo "ranks rc=$?"; tail -c 800 gpurun_out/r2m_bench_ranks_$N.json; tail -2 gpurun_out/r2m_bench_ranks_$N.err
