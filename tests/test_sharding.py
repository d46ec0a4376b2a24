"""Multi-rank host logic on the CPU (gloo, world_size 2): pictures shard independently, no data-path collective.
Each rank derives its shard with the same function bench.py uses, 'encodes' it with the CPU oracle restatement (the
GPU is not available here), and the ranks only exchange timing / counts -- exactly the N>1 plumbing of bench.py."""
import os
import subprocess
import sys
import textwrap

import numpy as np

import workloads as WL

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions():
    for n in (0, 1, 7, 64, 8192):
        for world in (1, 2, 3, 8):
            parts = [WL.shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1


def test_config3_generator_is_deterministic():
    a, b = WL.config3_image(5), WL.config3_image(5)
    assert a.shape == (512, 768) and a.dtype == np.uint8 and np.array_equal(a, b)
    assert not np.array_equal(a, WL.config3_image(6))


def test_two_rank_gloo_weak_scaling_plumbing(tmp_path):
    script = tmp_path / "rank.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys, time, hashlib
        sys.path.insert(0, {os.path.join(ROOT, 'tests')!r})
        import numpy as np, torch, torch.distributed as dist
        import workloads as WL, refutil as R
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        n_per_rank = 3
        lo = rank * n_per_rank                                  # weak scaling: every rank owns n_per_rank pictures
        imgs = [WL.config3_image(i)[:32, :64].copy() for i in range(lo, lo + n_per_rank)]
        dist.barrier(); t0 = time.perf_counter()
        outs = [R.oracle_encode(im, 2)[0] for im in imgs]        # stand-in for the GPU shard (no collective needed)
        dist.barrier(); dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)               # max over ranks, as bench.py does
        digest = hashlib.sha256(b"".join(outs)).digest()
        got = [None] * world
        dist.all_gather_object(got, (lo, n_per_rank, digest))
        if rank == 0:
            assert [g[0] for g in got] == [0, 3] and all(g[1] == 3 for g in got)
            assert got[0][2] != got[1][2]                        # different pictures per rank
            # the union equals a single-process run over all pictures
            ref = [R.oracle_encode(WL.config3_image(i)[:32, :64].copy(), 2)[0] for i in range(world * n_per_rank)]
            assert hashlib.sha256(b"".join(ref[:3])).digest() == got[0][2]
            assert hashlib.sha256(b"".join(ref[3:])).digest() == got[1][2]
            print("OK", float(dt))
        dist.destroy_process_group()
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29617", str(script)], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "OK" in r.stdout
