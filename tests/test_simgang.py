"""The gang-level logic of the kernel on the CPU: tests/sim/hevce_simgang.cpp runs one host thread per picture of a gang
(plus one per picture for team A of the 8x8 nodes) with real barriers, so what only exists with several pictures in a
CTA -- trial lanes packed across pictures, work handed to another picture's idle threads, the two teams' barrier
sequences -- is checked bit for bit against the oracle without a GPU, and under ThreadSanitizer for data races."""
import os
import shutil
import subprocess

import numpy as np
import pytest

import refutil as R
import simutil as S
import workloads as WL

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def crops(n, h=64, w=96):
    return [WL.config3_image(20 + i)[37 * i:37 * i + h, 41 * i:41 * i + w].copy() for i in range(n)]


def test_full_gang_mixed_qpd6():
    g = S.simgang().hevce_simgang_size()
    imgs, qs = crops(g), [i % 5 for i in range(g)]
    for order in (0, 3):
        for i, (s, r, err) in enumerate(S.simgang_encode(imgs, qs, order)):
            so, ro = R.oracle_encode(imgs[i], qs[i])
            assert err == 0 and s == so and np.array_equal(r, ro), (i, qs[i], order)


def test_short_gang_and_padding():
    imgs = crops(3, 45, 70)                      # padded to 64x96; four of the seven slots stay empty
    for i, (s, r, err) in enumerate(S.simgang_encode(imgs, [2, 4, 0])):
        so, ro = R.oracle_encode(imgs[i], [2, 4, 0][i])
        assert err == 0 and s == so and np.array_equal(r, ro), i


def test_gang_equals_single_picture_simulator():
    g = S.simgang().hevce_simgang_size()
    imgs = crops(g, 32, 64)
    for (s, r, err), img in zip(S.simgang_encode(imgs, [3] * g), imgs):
        s1, r1, e1 = S.sim_encode(img, 3)
        assert (s, err) == (s1, e1) and np.array_equal(r, r1)


def test_no_data_race_under_thread_sanitizer(tmp_path):
    """Positive control first (a deliberately racy program must be reported), then the gang simulator."""
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    racy = tmp_path / "racy.cpp"
    racy.write_text("#include <thread>\nint x;int main(){std::thread a([]{for(int i=0;i<100000;i++)x++;});"
                    "std::thread b([]{for(int i=0;i<100000;i++)x++;});a.join();b.join();return 0;}\n")
    if subprocess.run(["g++", "-O1", "-fsanitize=thread", "-pthread", "-o", str(tmp_path / "racy"), str(racy)],
                      capture_output=True).returncode != 0:
        pytest.skip("ThreadSanitizer not available")
    out = subprocess.run([str(tmp_path / "racy")], capture_output=True, text=True)
    if "ThreadSanitizer: data race" not in out.stderr:
        pytest.skip("ThreadSanitizer does not report races in this environment")
    main = tmp_path / "main.cpp"
    main.write_text(r'''
#include <cstdio>
#include <vector>
extern "C" int hevce_simgang_encode(int, unsigned char* const*, int, const unsigned char* const*, unsigned char* const*, int, int, const int*, int, int*, int*);
extern "C" int hevce_simgang_size();
int main() {
    const int n = hevce_simgang_size(), h = 32, w = 64, cap = 256 + 2 * h * w;
    std::vector<std::vector<unsigned char>> img(n, std::vector<unsigned char>(h * w)), out(n, std::vector<unsigned char>(cap)), rc(n, std::vector<unsigned char>(h * w));
    unsigned s = 12345;
    for (int i = 0; i < n; i++) for (int k = 0; k < h * w; k++) { s = s * 1664525u + 1013904223u; img[i][k] = (unsigned char)(((k % w) * 3 + (k / w) * 2 + (s >> 28)) & 255); }
    std::vector<unsigned char*> po(n), pr(n); std::vector<const unsigned char*> pi(n);
    std::vector<int> qs(n), lens(n), errs(n);
    for (int i = 0; i < n; i++) { po[i] = out[i].data(); pr[i] = rc[i].data(); pi[i] = img[i].data(); qs[i] = i % 5; }
    hevce_simgang_encode(n, po.data(), cap, pi.data(), pr.data(), h, w, qs.data(), 0, lens.data(), errs.data());
    for (int i = 0; i < n; i++) printf("%d %d %d\n", i, lens[i], errs[i]);
}
''')
    exe = tmp_path / "tsan_gang"
    subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=thread", "-pthread", "-I", S.CSRC, "-o", str(exe), str(main),
                    os.path.join(S.SIM_DIR, "hevce_simgang.cpp")], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=1200)
    assert out.returncode == 0 and "ThreadSanitizer" not in out.stderr, out.stderr[-3000:]
    rows = [l.split() for l in out.stdout.strip().splitlines()]
    assert len(rows) == S.simgang().hevce_simgang_size() and all(int(r[1]) > 100 and int(r[2]) == 0 for r in rows)


@pytest.mark.parametrize("variant", ["g4", "g2", "w1"])
def test_other_variants_full_and_short_gangs(variant):
    """The 4- and 2-picture gangs and the wide one-picture variant (other thread counts per picture, other team sizes,
    few trial lanes per warp, the large pool plan of the wide variant) through the same source, incl. a short gang."""
    g = S.simgang(variant).hevce_simgang_size()
    for n in sorted({g, max(1, g - 1)}):
        imgs, qs = crops(n, 45, 70), [(2 * i + 1) % 5 for i in range(n)]
        for i, (s, r, err) in enumerate(S.simgang_encode(imgs, qs, 0, variant=variant)):
            so, ro = R.oracle_encode(imgs[i], qs[i])
            assert err == 0 and s == so and np.array_equal(r, ro), (variant, n, i)


def test_wide_variant_single_thread_simulator_permuted():
    """The wide pool plan (all 35 candidates of a 16x16 / 32x32 step in one round) under permuted work-item orders."""
    img = WL.config3_image(7)[100:164, 300:396].copy()
    for q, order in ((0, 1), (2, 3), (4, 5)):
        s, r, err = S.sim_encode(img, q, order, variant="w1")
        so, ro = R.oracle_encode(img, q)
        assert err == 0 and s == so and np.array_equal(r, ro), (q, order)


@pytest.mark.parametrize("variant", ["t1", "c2"])
def test_track_simulator_matches_oracle(variant):
    """Parent || child variants (t1: three tracks in one CTA; c2: the tracks on a two-CTA cluster, picture state pushed into
    the parent tracks' blocks, results fetched back): one host thread per track, real rendezvous; a 64x96 crop and a padded
    45x70 picture at every qpd6 against the oracle."""
    for k, (h, w) in enumerate(((64, 96), (45, 70))):
        img = crops(1, h, w)[0]
        for q in range(5):
            s, r, err = S.simtrack_encode(img, q, order=(0, 3)[(q + k) % 2], variant=variant)
            so, ro = R.oracle_encode(img, q)
            assert err == 0 and s == so and np.array_equal(r, ro), (variant, h, w, q)
            s1, r1, e1 = S.sim_encode(img, q, variant=variant)    # the same variant with its tracks one after the other
            assert (s1, e1) == (s, err) and np.array_equal(r1, r)


@pytest.mark.parametrize("variant", ["t1", "c2"])
def test_tracks_share_no_unsynchronised_data_under_thread_sanitizer(tmp_path, variant):
    """ThreadSanitizer over the track simulator: the 16x16 / 32x32 candidate tracks run beside the 8x8 chain."""
    probe = tmp_path / "probe.cpp"
    probe.write_text("#include <thread>\nint x;int main(){std::thread a([]{for(int i=0;i<100000;i++)x++;});"
                     "std::thread b([]{for(int i=0;i<100000;i++)x++;});a.join();b.join();return 0;}\n")
    if subprocess.run(["g++", "-O1", "-fsanitize=thread", "-pthread", "-o", str(tmp_path / "probe"), str(probe)],
                      capture_output=True).returncode != 0:
        pytest.skip("ThreadSanitizer not available")
    if "ThreadSanitizer: data race" not in subprocess.run([str(tmp_path / "probe")], capture_output=True, text=True).stderr:
        pytest.skip("ThreadSanitizer does not report races in this environment")
    main = tmp_path / "main.cpp"
    main.write_text(r'''
#include <cstdio>
#include <vector>
extern "C" int hevce_simtrack_encode(unsigned char*, int, const unsigned char*, unsigned char*, int*, int*, int, int, int*);
int main() {
    const int h = 64, w = 96, cap = 256 + 2 * h * w;
    std::vector<unsigned char> img(h * w), out(cap), rc(h * w);
    unsigned s = 777;
    for (int k = 0; k < h * w; k++) { s = s * 1664525u + 1013904223u; img[k] = (unsigned char)(((k % w) * 2 + (k / w) * 3 + (s >> 27)) & 255); }
    for (int q = 0; q < 5; q += 2) {
        int ys = h, xs = w, err = 0;
        int n = hevce_simtrack_encode(out.data(), cap, img.data(), rc.data(), &ys, &xs, q, 0, &err);
        printf("%d %d %d\n", q, n, err);
    }
}
''')
    exe = tmp_path / "tsan_tracks"
    subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=thread", "-pthread", *S.variant_flags(variant), "-I", S.CSRC,
                    "-o", str(exe), str(main), os.path.join(S.SIM_DIR, "hevce_simtrack.cpp")], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=1200)
    assert out.returncode == 0 and "ThreadSanitizer" not in out.stderr, out.stderr[-3000:]
    rows = [l.split() for l in out.stdout.strip().splitlines()]
    assert len(rows) == 3 and all(int(r[1]) > 100 and int(r[2]) == 0 for r in rows)
