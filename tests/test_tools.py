"""Host-side helpers of SURVEY.md section 8f row f4 (tools/hevc_eval.py, tools/convert_to_pgm.py): the parts that need
no device."""
import os
import sys

import numpy as np
import pytest

import refutil as R

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import convert_to_pgm as C   # noqa: E402
import hevc_eval as E        # noqa: E402


def test_image_pad_is_edge_replication():
    a = np.arange(45 * 70, dtype=np.uint8).reshape(45, 70)
    p = E.image_pad(a)
    assert p.shape == (64, 96) and np.array_equal(p[:45, :70], a)
    assert (p[:45, 70:] == a[:, -1:]).all() and (p[45:, :70] == a[-1:, :]).all() and (p[45:, 70:] == a[-1, -1]).all()
    assert E.image_pad(p) is not None and E.image_pad(p).shape == p.shape


def test_ssim_properties():
    rng = np.random.default_rng(0)
    a = np.clip(np.cumsum(rng.integers(-3, 4, (64, 80)), axis=1) + 128, 0, 255).astype(np.uint8)
    assert E.ssim(a, a) == pytest.approx(1.0, abs=1e-12)
    n1 = np.clip(a.astype(int) + rng.integers(-4, 5, a.shape), 0, 255).astype(np.uint8)
    n2 = np.clip(a.astype(int) + rng.integers(-20, 21, a.shape), 0, 255).astype(np.uint8)
    s1, s2 = E.ssim(a, n1), E.ssim(a, n2)
    assert 1.0 > s1 > s2 > 0.0
    assert E.ssim(a, n1) == pytest.approx(E.ssim(n1, a), abs=1e-12)
    flat = np.full((32, 32), 9, np.uint8)            # constant pictures: luminance term only
    assert E.ssim(flat, flat) == pytest.approx(1.0)
    c1 = (0.01 * 256) ** 2
    assert E.ssim(flat, flat + 100) == pytest.approx((2 * 9 * 109 + c1) / (81 + 109 * 109 + c1), rel=1e-9)


def test_convert_to_pgm_round_trip(tmp_path):
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(2)
    rgb = rng.integers(0, 256, (37, 53, 3)).astype(np.uint8)
    src = tmp_path / "in"; dst = tmp_path / "out"
    src.mkdir()
    Image.fromarray(rgb).save(src / "a.png")
    (src / "junk.txt").write_text("not a picture")
    assert C.main(["x", str(src), str(dst)]) == 0
    assert sorted(os.listdir(dst)) == ["a.pgm"]
    got = R.read_pgm(str(dst / "a.pgm"))
    assert np.array_equal(got, np.asarray(Image.fromarray(rgb).convert("L")))
    assert C.main(["x", str(src / "a.png"), str(tmp_path / "single")]) == 0 and os.path.exists(tmp_path / "single.pgm")
    assert C.main(["x"]) == -1


def test_match_ssim_bisection(tmp_path):
    pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(4)
    yy, xx = np.mgrid[0:96, 0:128]
    img = np.clip(100 + 40 * np.sin(xx / 9.0) + 30 * np.cos(yy / 7.0) + rng.integers(-6, 7, (96, 128)), 0, 255).astype(np.uint8)
    p = str(tmp_path / "t.jpg")
    E.save_as(img, p, 60)
    target = E.ssim(img, E.read_monochrome(p))
    s, size, q = E.match_ssim(img, p, target, 1, 101)
    assert abs(q - 60) <= 1 and abs(s - target) < 5e-3 and size == os.path.getsize(p)
