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
