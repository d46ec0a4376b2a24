#!/bin/bash
mkdir -p gpurun_out
timeout 120 python - <<'PY' 2>&1 | tee gpurun_out/r2r_c2_first.log
import sys, time
sys.path.insert(0, "hevc-image-encoder-lite_b200"); sys.path.insert(0, "tests")
import numpy as np, hevce_b200 as H, golden_util as G
H.set_variant("c2")
data, _ = G.small_cases()
for name, q in (("k01_32x32", 2), ("k01_64x64", 2), ("k01_45x70", 0), ("noise_64", 4)):
    t = time.time()
    s, r = H.HEVCImageEncoder(np.array(data[f"{name}/in"]), q)
    ok = s == data[f"{name}/q{q}/stream"].tobytes() and np.array_equal(r, data[f"{name}/q{q}/rcon"])
    print(name, q, "OK" if ok else "MISMATCH", f"{time.time() - t:.2f}s", flush=True)
PY
echo "first rc=$?"
timeout 300 python -m pytest tests/test_gpu_variants.py -q -x -k "c2 or agree or automatic" 2>&1 | tail -4 | tee -a gpurun_out/r2r_c2_first.log
timeout 200 python tools/variant_bench.py 64 64 2 w1 t1 c2 2>&1 | tee gpurun_out/r2r_variants.log
timeout 200 python tools/variant_bench.py 64 64 4 w1 t1 c2 2>&1 | tee -a gpurun_out/r2r_variants.log
