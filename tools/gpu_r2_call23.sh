#!/bin/bash
# A/B of the coder / RDOQ diet: the library of the previous commit (0old), the head (base = BINTAB=1), explicit ld.shared
# for the bin table (BINTAB=2), predicated byte release, three-candidate RDOQ; then the parity tests on the head.
mkdir -p gpurun_out
python tools/ab_variants.py run g7 > gpurun_out/r2aa_ab_g7.log 2>&1; cat gpurun_out/r2aa_ab_g7.log
python tools/ab_variants.py run c2 > gpurun_out/r2aa_ab_c2.log 2>&1; cat gpurun_out/r2aa_ab_c2.log
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x > gpurun_out/r2aa_parity.log 2>&1; tail -3 gpurun_out/r2aa_parity.log
