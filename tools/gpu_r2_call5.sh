#!/bin/bash
mkdir -p gpurun_out
for v in w1 t1 g7; do
n=148; [ $v = g7 ] && n=1036
HEVCE_VARIANT=$v python tools/phase_profile.py $n 64 64 2 > gpurun_out/r2e_phases_$v.log 2>&1; cat gpurun_out/r2e_phases_$v.log
done
