"""Decoder-based conformance, CPU part: the streams the kernel source produces (through
the host-compiled simulator) are decoded by libavcodec and must equal deblock(img_rcon) -- the reconstruction passed
through the in-loop deblocking filter that the reference's stream header leaves enabled -- at every qpd6."""
import numpy as np
import pytest

import deblock_model as D
import decode_util as U
import simutil as S
import workloads as WL


def pictures():
    rng = np.random.default_rng(5)
    yy, xx = np.mgrid[0:96, 0:128]
    return {
        "gradient": np.clip(40 + yy * 1.1 + xx * 0.7 + rng.integers(-1, 2, (96, 128)), 0, 255).astype(np.uint8),   # 32x32 CUs, strong filter
        "steps": (np.clip((xx // 16) * 12 + (yy // 16) * 9 + 30, 0, 255) + rng.integers(0, 2, (96, 128))).astype(np.uint8),
        "noise": rng.integers(0, 256, (64, 64)).astype(np.uint8),
        "kodak": WL.config3_image(7)[300:396, 100:228],
        "ragged": WL.config3_image(2)[10:55, 20:90],                                                                # padded 45x70
    }


def test_tables():
    assert D.BETA[16] == 6 and D.BETA[28] == 18 and D.BETA[51] == 64
    assert D.TC[17] == 0 and D.TC[18] == 1 and D.TC[30] == 2 and D.TC[53] == 24


def test_identity_at_low_qp():
    r = np.random.default_rng(1).integers(0, 256, (64, 64)).astype(np.uint8)
    cu, kind = np.full((16, 16), 8, np.uint8), np.zeros((8, 8), np.uint8)
    for q in (0, 1):
        assert np.array_equal(D.deblock(r, cu, kind, q), r)


def test_interior_edges_of_large_transform_blocks_stay():
    """A 32x32 CU with one TU has no transform edge inside: only its outline may change."""
    r = np.zeros((64, 64), np.uint8)
    r[:, 32:] = 6                                  # a small step on the CU boundary, and one inside a CU at x = 16
    r[:, 16:32] = 3
    cu, kind = np.full((16, 16), 32, np.uint8), np.zeros((8, 8), np.uint8)
    out = D.deblock(r, cu, kind, 4)
    assert np.array_equal(out[:, 8:24], r[:, 8:24])          # the inner step at x = 16 is not on a TU edge
    assert not np.array_equal(out[:, 28:36], r[:, 28:36])    # the CU boundary at x = 32 is smoothed


@pytest.mark.parametrize("q", [0, 1, 2, 3, 4])
def test_decoder_equals_deblocked_reconstruction(q):
    n = 0
    for name, img in pictures().items():
        s, r, err = S.sim_encode(img, q)
        assert err == 0
        luma = U.decode_luma(s, r.shape)
        if luma is None:
            pytest.skip("no HEVC decoder in this OpenCV build")
        cu, mode, kind = S.sim_last_partition(r.shape)
        assert set(np.unique(cu)) <= {8, 16, 32} and kind.max() <= 2 and mode.max() <= 34
        assert np.array_equal(luma, D.deblock(r, cu, kind, q)), (name, q)
        n += 1
    assert n == 5
