"""GPU tests of the two "next" rows of SURVEY.md section 8f that live on the device side:
  f2 -- decoder-based conformance: libavcodec decodes the streams libhevce_b200.so produced and must output
        deblock(img_rcon) (tests/deblock_model.py) at every qpd6, including pictures that never had an oracle run;
  f3 -- per-picture MSE/PSNR reduced on the device (hevce_session_quality) against calcImagePSNR's arithmetic.
Everything goes through the C ABI.  Nothing here reads /root/reference."""
import os

import numpy as np
import pytest

import deblock_model as D
import decode_util as U
import golden_util as G
import workloads as WL

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def H():
    import hevce_b200
    assert os.path.exists(hevce_b200.LIB_PATH), "libhevce_b200.so missing: the CUDA extension must be built in-tree"
    return hevce_b200


def ref_quality(img, rcon):
    """calcImagePSNR (HEVCeMain.c:116-133) in numpy."""
    hm, wm = min(img.shape[0], rcon.shape[0]), min(img.shape[1], rcon.shape[1])
    d = img[:hm, :wm].astype(np.int64) - rcon[:hm, :wm].astype(np.int64)
    mse = max(float((d * d).sum()) / hm / wm, 1e-9)
    return mse, 10.0 * np.log10(255 * 255 / mse)


def test_quality_matches_reference_arithmetic(H):
    data, _ = G.small_cases()
    imgs = [data[f"{n}/in"] for n in G.small_case_names()]
    imgs += [WL.config3_image(5), WL.config3_image(6)[:500, :701], np.full((40, 40), 77, np.uint8)]
    qs = [i % 5 for i in range(len(imgs))]
    ses = H.Session(0, [i.shape for i in imgs], qs)
    ses.upload(imgs)
    ses.encode()
    mse, psnr = ses.quality()
    assert ses.quality_ms > 0
    _, rcons = ses.download()
    ses.close()
    for i, (img, r) in enumerate(zip(imgs, rcons)):
        m, p = ref_quality(img, r)
        assert mse[i] == pytest.approx(m, rel=1e-12, abs=0), i      # exact integer SSE, one double division
        assert psnr[i] == pytest.approx(p, rel=1e-12), i
    lossless = [i for i in range(len(imgs)) if mse[i] == 1e-9]      # the MSE floor (flat pictures reconstruct exactly)
    assert lossless, "expected at least one picture at the 1e-9 floor"


def test_quality_of_clamped_picture(H):
    """A picture taller than the 8192 clamp: quality is taken over the rows both buffers have (HEVCeMain.c:117-118)."""
    rng = np.random.default_rng(3)
    img = np.clip(np.cumsum(rng.integers(-2, 3, (8200, 33)), axis=0) + 120, 0, 255).astype(np.uint8)
    ses = H.Session(0, [img.shape], 3)
    ses.upload([img])
    ses.encode()
    mse, psnr = ses.quality()
    _, rcons = ses.download()
    ses.close()
    assert rcons[0].shape == (8192, 64)
    m, p = ref_quality(img, rcons[0])
    assert mse[0] == pytest.approx(m, rel=1e-12) and psnr[0] == pytest.approx(p, rel=1e-12)


def test_partition_maps_are_consistent(H):
    img = WL.config3_image(9)
    ses = H.Session(0, [img.shape], 2)
    ses.upload([img])
    ses.encode()
    cu, mode, kind = ses.partition(0)
    ses.close()
    assert cu.shape == (128, 192) and kind.shape == (64, 96)
    assert set(np.unique(cu)) <= {8, 16, 32} and mode.max() <= 34 and kind.max() <= 2
    for s in (8, 16, 32):                                           # CUs are aligned squares of one size, mode and kind
        u = s // 4
        for y in range(0, 128, u):
            for x in range(0, 192, u):
                if cu[y, x] == s and y % u == 0 and x % u == 0:
                    assert (cu[y:y + u, x:x + u] == s).all()
                    k = kind[y // 2, x // 2]
                    assert (kind[y // 2:(y + u + 1) // 2, x // 2:(x + u + 1) // 2] == k).all()
                    if k != 2:
                        assert (mode[y:y + u, x:x + u] == mode[y, x]).all()
    assert (kind[cu[::2, ::2] > 8] != 2).all()                      # NxN exists at 8x8 only


@pytest.mark.parametrize("q", [0, 1, 2, 3, 4])
def test_decoder_equals_deblocked_reconstruction(H, q):
    """Full-size pictures, including config-3 indices far beyond anything the oracle was ever run on."""
    K = WL.kodak_landscape()
    imgs = [K[0], K[12], K[22], WL.config3_image(4001, K), WL.config3_image(7777, K)[:301, :455]]
    ses = H.Session(0, [i.shape for i in imgs], q)
    ses.upload(imgs)
    ses.encode()
    streams, rcons = ses.download()
    parts = [ses.partition(i) for i in range(len(imgs))]
    ses.close()
    for i, (s, r) in enumerate(zip(streams, rcons)):
        luma = U.decode_luma(s, r.shape)
        if luma is None:
            pytest.skip("no HEVC decoder in this OpenCV build")
        cu, mode, kind = parts[i]
        assert np.array_equal(luma, D.deblock(r, cu, kind, q)), (i, q)


def test_quality_study_on_the_batch_api(H, tmp_path):
    """f4: tools/hevc_eval.py on three Kodak crops; the HEVC column must be the batch encoder's bytes and, as the
    reference's README reports for the full set, HEVC needs fewer bits than JPEG at equal SSIM."""
    pytest.importorskip("PIL.Image")
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import hevc_eval as E
    from PIL import Image
    K = WL.kodak_landscape()
    src = tmp_path / "in"
    src.mkdir()
    crops = [K[3][64:64 + 250, 100:100 + 375], K[14][:256, :384], K[20][200:456, 300:684]]
    for i, c in enumerate(crops):
        Image.fromarray(c).save(src / f"p{i}.png")
    rows, means = E.evaluate(str(src), str(tmp_path / "out"), 3, formats=E.COMPARISONS[:1] + E.COMPARISONS[2:], log=lambda *_: None)
    assert [r["name"] for r in rows] == ["p0.png", "p1.png", "p2.png"]
    for r, c in zip(rows, crops):
        s, rc = H.HEVCImageEncoder(E.image_pad(c), 3)
        base = str(tmp_path / "out" / os.path.splitext(r["name"])[0])
        assert open(base + ".h265", "rb").read() == s
        assert r["HEVC"][1] == pytest.approx(8.0 * len(s) / rc.size)
        assert abs(r["JPEG"][0] - r["HEVC"][0]) < 0.02 and os.path.exists(base + ".jpg") and os.path.exists(base + ".webp")
    assert means["JPEG"] > means["HEVC"]


def test_quality_and_partition_need_an_encode(H):
    img = np.full((40, 40), 9, np.uint8)
    ses = H.Session(0, [img.shape], 1)
    ses.upload([img])
    with pytest.raises(H.HevceError) as e1:
        ses.quality()
    with pytest.raises(H.HevceError) as e2:
        ses.partition(0)
    assert e1.value.code == H.ERR_STATE and e2.value.code == H.ERR_STATE
    ses.encode()
    mse, _ = ses.quality()
    assert mse[0] >= 1e-9
    ses.close()
