#!/bin/bash
# A/B #4: software-pipelined context load in the sig-flag loop (CTXFWD=2), branch-free byte release inside the release branch (FLUSH=1), short-cut release (FASTREL=1).
mkdir -p gpurun_out
python tools/ab_variants.py run g7 > gpurun_out/r2ad_ab_g7.log 2>&1; cat gpurun_out/r2ad_ab_g7.log
HEVCE_AB_DIR=$PWD/hevc-image-encoder-lite_b200/ab_c2 python tools/ab_variants.py run c2 > gpurun_out/r2ad_ab_c2.log 2>&1; cat gpurun_out/r2ad_ab_c2.log
