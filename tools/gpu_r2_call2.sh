#!/bin/bash
# first run of the multi-variant library: parity suite, per-variant latency, wide mode on a 4K picture, bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2b_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_gputests.log
tail -5 gpurun_out/r2b_gputests.log
python tools/variant_bench.py 64 64 2 > gpurun_out/r2b_variants.log 2>&1; cat gpurun_out/r2b_variants.log
for v in g7 g4 g2 w1; do
python - $v <<'PY' 2>&1 | tee -a gpurun_out/r2b_ragged.log
import sys, time
sys.path.insert(0, "hevc-image-encoder-lite_b200"); sys.path.insert(0, "tests")
import numpy as np, hevce_b200 as H, golden_util as G
v = sys.argv[1]
H.set_variant(v)
data, _ = G.small_cases()
imgs, qs, keys = [], [], []
for n in G.small_case_names():
    for q in range(5):
        imgs.append(data[f"{n}/in"]); qs.append(q); keys.append((n, q))
t = time.time()
streams, rcons = H.HEVCImageEncoderBatch(imgs, qs)
t = time.time() - t
bad = [(n, q) for (n, q), s, r in zip(keys, streams, rcons) if s != data[f"{n}/q{q}/stream"].tobytes() or not np.array_equal(r, data[f"{n}/q{q}/rcon"])]
print(v, "ragged 85:", "OK" if not bad else f"BAD {bad[:5]}", f"{t:.2f}s")
PY
done
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -c 1500 gpurun_out/r2b_bench.json; tail -3 gpurun_out/r2b_bench.err
