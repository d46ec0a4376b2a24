#!/usr/bin/env python
"""bench.py -- Mpixel/s encoded (bit-exact bitstream) on Kodak-size batches, 1..8 B200, next to the reference's CPU build.

    python bench.py --gpus N --steps K --warmup W              # our arm (N>1: launched by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's own CPU implementation

A "step" is one pass of the hot path over one batch: `--images` (default 1036 = 148 SMs x 7 pictures per CTA) synthetic
768x512 pictures PER GPU at qpd6=2 -- BASELINE.json configs[2] sharded (weak scaling: per-GPU work fixed; pictures are
independent, no collective on the data path, NCCL only carries the timing barrier / max-reduce).
  value : whole-job Mpixel/s with the inputs already resident in HBM (session upload outside the timed region)
  e2e   : the same metric through the public C entry point HEVCImageEncoderBatch with HOST buffers (H2D of the
          pictures, D2H of streams + reconstructions inside the timed region)
Every timed run is parity-gated: the streams of the timed batch must hash identically on every step, and a sample is
compared byte-for-byte with the CPU oracle (oracle/_ref when present, else the restatement).
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "hevc-image-encoder-lite_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

W_INT_OPS_PER_PIXEL = 9782          # minimal (partial-butterfly) transform arithmetic, exact
KODAK_PIXELS = 768 * 512


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], threading.Event()
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(int(float(r[0])) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(float(self.rows[0][1])), "reasons": reasons, "samples": len(sm)}


def cpu_checker():
    import refutil as R
    if os.path.exists(R.REF_SO):
        return R.ref(), "reference"
    return R.oracle(), "port"


def cpu_sample_run(imgs, q, threads):
    """Encode `imgs` with the CPU checker on `threads` host threads (ctypes releases the GIL; the reference is
    re-entrant, README.md:25-28). Returns seconds."""
    import refutil as R
    lib, _ = cpu_checker()
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(lambda im: R.encode_with(lib, im, q), imgs))
    return time.perf_counter() - t0


def run_reference(a, rank, world):
    import workloads as WL
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    _, kind = cpu_checker()
    imgs = WL.config3_batch(0, cores)
    times = []
    for s in range(a.warmup + a.steps):
        dt = cpu_sample_run(imgs, a.qpd6, cores)
        if s >= a.warmup:
            times.append(dt)
    t = sum(times)
    mpx = len(imgs) * KODAK_PIXELS * len(times) / t / 1e6
    line = {
        "impl": "reference", "metric": "Mpixel/s encoded, bit-exact bitstream, Kodak-size batch", "value": mpx, "unit": "Mpixel/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * t / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": "configs[2]: synthetic 768x512 8-bit grayscale (Kodak-derived), qpd6=2", "qpd6": a.qpd6},
        "cpu_baseline": {"value": mpx, "unit": "Mpixel/s", "cores": cores, "kind": kind,
                         "sample": f"{len(imgs)} pictures of the workload per step, one picture per host thread"},
        "e2e": {"value": mpx, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--images", type=int, default=1036, help="pictures per GPU per step (1036 = 148 SMs x 7 pictures per CTA: one full wave)")
    ap.add_argument("--qpd6", type=int, default=2)
    ap.add_argument("--cpu-sample", type=int, default=-1, help="pictures for the cpu_baseline leg (-1 = one per host core)")
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        run_reference(a, rank, world)
        return

    import torch
    import hevce_b200 as H
    import refutil as R
    import workloads as WL
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the encoder has no CPU path")
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if not dist:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    H.set_devices([local])
    n = a.images
    lo = rank * n                                      # weak scaling: rank r encodes pictures [r*n, (r+1)*n)
    imgs = WL.config3_batch(lo, n)
    shapes = [i.shape for i in imgs]
    pixels_step = n * KODAK_PIXELS

    # ---- device-resident arm: `value`
    ses = H.Session(local, shapes, a.qpd6)
    ses.upload(imgs)
    for _ in range(a.warmup):
        ses.encode()
    launches0 = ses.launches
    kernel_ms, commit_ms = [], []
    barrier()
    with ClockSampler(local) as clk:
        t0 = time.perf_counter()
        for _ in range(a.steps):
            kernel_ms.append(ses.encode())             # CUDA events on the launching stream, inside the library
            commit_ms.append(ses.commit_ms)
        barrier()
        dt = time.perf_counter() - t0
    dt = max_over_ranks(dt)
    launches = ses.launches - launches0
    streams, rcons = ses.download()
    digest = hashlib.sha256(b"".join(hashlib.sha256(s).digest() for s in streams)).hexdigest()
    grid = ses.grid
    ses.close()
    value = world * pixels_step * a.steps / dt / 1e6

    # ---- parity gate on the timed batch (sample vs the CPU checker)
    lib, kind = cpu_checker()
    parity_n = 1 if a.no_cpu else 2
    for k in range(parity_n):
        ws, wr = R.encode_with(lib, imgs[k], a.qpd6)
        if ws != streams[k] or not np.array_equal(wr, rcons[k]):
            raise SystemExit(f"bench.py: PARITY FAILURE on picture {lo + k}: timed output differs from the CPU {kind}")

    # ---- end-to-end arm through the public C entry point with host buffers
    e2e_steps = a.steps
    host_out = H.alloc_outputs(shapes)                     # caller-owned pbuffer / img_rcon arrays, reused like a C caller would
    H.HEVCImageEncoderBatch(imgs, a.qpd6, outputs=host_out, copy_streams=False)   # warm-up: pooled session, pinned staging, page faults
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        s2, r2 = H.HEVCImageEncoderBatch(imgs, a.qpd6, outputs=host_out, copy_streams=False)
    barrier()
    dt2 = max_over_ranks(time.perf_counter() - t0)
    if hashlib.sha256(b"".join(hashlib.sha256(bytes(s)).digest() for s in s2)).hexdigest() != digest:
        raise SystemExit("bench.py: e2e arm produced different streams than the device-resident arm")
    e2e_value = world * pixels_step * e2e_steps / dt2 / 1e6
    h2d = sum(i.size for i in imgs)
    d2h = sum(r.size for r in r2) + sum(len(s) for s in s2) + 8 * n

    # ---- roofline of the (only) kernel: integer issue, SURVEY.md section 8d
    km = sum(kernel_ms) / len(kernel_ms) * 1e-3
    int_peak = H.measure_int_peak(local)
    achieved = pixels_step * W_INT_OPS_PER_PIXEL / km
    pk, pk_src = peaks()
    stream_bytes = sum(len(s) for s in streams)
    hbm_bytes = 2 * pixels_step + stream_bytes           # 1 B/px read + 1 B/px recon written + bitstream
    roofline = {
        "bound": "int_issue", "kernel": "hevce_encode_kernel",
        "achieved": achieved / 1e12, "peak": int_peak / 1e12, "unit": "Tint-op/s", "frac": achieved / int_peak,
        "peak_source": "measured live: hevce_int_peak_kernel (IMAD=2 ops + LOP3 + IADD3 chains); MEASURED_PEAKS.json has no integer figure",
        "work_per_pixel": W_INT_OPS_PER_PIXEL, "kernel_ms": km * 1e3, "commit_kernel_ms": sum(commit_ms) / len(commit_ms), "traffic": None,
        "hbm": {"algorithmic_bytes": hbm_bytes, "achieved_gbs": hbm_bytes / km / 1e9, "peak_gbs": pk["hbm_gbs"],
                "frac": hbm_bytes / km / 1e9 / pk["hbm_gbs"], "peak_source": pk_src},
    }

    line = {
        "metric": "Mpixel/s encoded, bit-exact bitstream, Kodak-size batch", "value": value, "unit": "Mpixel/s",
        "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * dt / a.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": "configs[2] shard: synthetic 768x512 8-bit grayscale (Kodak-derived, tests/workloads.py), qpd6=2",
                   "images_per_gpu": n, "qpd6": a.qpd6, "pixels_per_step_per_gpu": pixels_step, "grid_ctas": grid,
                   "l2": f"inputs larger than L2: {n * KODAK_PIXELS / 1e6:.0f} MB of pictures + {n * 384 * 2.45e-3:.0f} MB of per-CTU records/levels written and re-read per step",
                   "parallelism": f"{world} x independent shards, no collective"},
        "e2e": {"value": e2e_value, "unit": "Mpixel/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "HEVCImageEncoderBatch (host buffers, pinned staging inside the library)", "steps": e2e_steps},
        "gpu_launches": launches * world,
        "clocks": clk.summary(),
        "roofline": roofline,
        "parity": {"checked_pictures": parity_n, "against": kind, "stream_digest": digest[:16]},
    }

    if rank == 0 and not a.no_cpu:
        cores = os.cpu_count() or 1
        m = cores if a.cpu_sample < 0 else a.cpu_sample
        sample = imgs[:m] if m <= n else WL.config3_batch(lo, m)
        t = cpu_sample_run(sample, a.qpd6, cores)
        line["cpu_baseline"] = {"value": len(sample) * KODAK_PIXELS / t / 1e6, "unit": "Mpixel/s", "cores": cores, "kind": kind,
                                "sample": f"first {len(sample)} pictures of the timed batch, one per host thread, {t:.1f} s"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
