"""Kernel logic on the CPU: csrc/hevce_core.h compiled for the host as a single-threaded CTA simulator
(tests/sim, test infrastructure only) must reproduce the reference byte for byte, in any work-item order."""
import numpy as np
import pytest

import golden_util as G
import simutil as S


@pytest.mark.parametrize("name", G.small_case_names())
def test_sim_matches_reference_small(name):
    data, _ = G.small_cases()
    img = data[f"{name}/in"]
    for q in range(5):
        order = (0, 1, 3, 5, 9)[(q + len(name)) % 5]       # permuted PAR_FOR item order: phases must be race-free
        s, r, err = S.sim_encode(img, q, order)
        assert err == 0
        assert s == data[f"{name}/q{q}/stream"].tobytes(), (name, q, order)
        assert np.array_equal(r, data[f"{name}/q{q}/rcon"]), (name, q, order)


@pytest.mark.slow
def test_sim_kodak_01_q2():
    imgs, man = G.kodak()
    import refutil as R
    s, r, err = S.sim_encode(imgs["k01"], 2)
    assert err == 0 and R.sha(s) == man["01"]["q"]["2"]["stream_sha256"] and R.sha(r.tobytes()) == man["01"]["q"]["2"]["rcon_sha256"]


def _kodak_job(args):
    img, q = args
    s, r, err = S.sim_encode(img, q)
    import refutil as R
    return len(s), R.sha(s), R.sha(r.tobytes()), err


@pytest.mark.slow
def test_sim_kodak_all_120():
    """All 24 Kodak pictures x qpd6 0..4 through the kernel source on the CPU against the reference manifest
    (~3 minutes on 8 cores; opt-in with HEVCE_SLOW=1).  The same 120 encodes run on the GPU in test_gpu_parity.py."""
    import os
    from multiprocessing import Pool
    z, man = G.kodak()
    imgs = {j: np.array(z[f"k{j:02d}"]) for j in range(1, 25)}   # materialised before the fork: the npz handle is not fork-safe
    S.build_sim()
    keys = [(j, q) for j in range(1, 25) for q in range(5)]
    with Pool(min(16, os.cpu_count() or 1)) as pool:
        out = pool.map(_kodak_job, [(imgs[j], q) for j, q in keys], chunksize=1)
    for (j, q), (n, hs, hr, err) in zip(keys, out):
        m = man[f"{j:02d}"]["q"][str(q)]
        assert err == 0 and n == m["len"] and hs == m["stream_sha256"] and hr == m["rcon_sha256"], (j, q)
