"""HEVC luma deblocking filter (ITU-T H.265 section 8.7.2) restricted to what the reference's streams use -- TEST
INFRASTRUCTURE ONLY.

The stream leaves the in-loop filter enabled (slice_deblocking_filter_disabled_flag = 0, offsets 0, no SAO; header
bytes at HEVCe.c:665-691), while `img_rcon` is the encoder's unfiltered reconstruction.  A conforming decoder
therefore outputs deblock(img_rcon).  This model lets the conformance tests assert `decode(stream) ==
deblock(img_rcon)` at every qpd6; at qpd6 0-1 both beta and tc are 0 and the filter
is the identity.

Inputs: the reconstruction, the CU size per 4x4 unit and the CU kind per 8x8 unit (0 one TU, 1 four TUs, 2 NxN) as
returned by hevce_session_partition / the simulator, and qpd6 (QP = 6*qpd6 + 4, uniform over the picture).
All CUs are intra, so every transform/prediction edge on the 8x8 grid has boundary strength 2.
"""
import numpy as np

BETA = [0] * 16 + [6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18] + list(range(20, 66, 2))            # Table 8-12, Q = 0..51
TC = [0] * 18 + [1] * 9 + [2] * 4 + [3] * 4 + [4] * 3 + [5, 5, 6, 6, 7, 8, 9, 10, 11, 13, 14, 16, 18, 20, 22, 24]  # Q = 0..53
assert len(BETA) == 52 and len(TC) == 54


def tu_size_map(cu_size, kind):
    """Transform-block size per 4x4 unit: CU size for kind 0, half of it for kind 1; an NxN CU has four 4x4 TUs."""
    k4 = np.repeat(np.repeat(kind, 2, axis=0), 2, axis=1)
    cu = cu_size.astype(np.int32)
    return np.where(k4 == 0, cu, np.where(k4 == 1, cu // 2, 4))


def _filter_edges(pic, edge, beta, tc):
    """Filter all vertical edges of `pic` (int32, HxW).  edge[y4, x8]: the edge at x = 8*x8 is filtered on the 4-row
    segment y4.  Returns the filtered picture; horizontal edges are handled by the caller through a transpose."""
    h, w = pic.shape
    out = pic.copy()
    xs = np.arange(8, w, 8)
    if len(xs) == 0:
        return out
    # samples across every candidate edge: P[..., i] = p_i, Q[..., i] = q_i, shape (H, nx, 4)
    P = np.stack([pic[:, xs - 1 - i] for i in range(4)], axis=-1)
    Q = np.stack([pic[:, xs + i] for i in range(4)], axis=-1)
    seg = lambda a: a.reshape(h // 4, 4, *a.shape[1:])          # (seg, line, nx, ...)
    Ps, Qs = seg(P), seg(Q)
    dp = np.abs(Ps[..., 2] - 2 * Ps[..., 1] + Ps[..., 0])       # (seg, line, nx)
    dq = np.abs(Qs[..., 2] - 2 * Qs[..., 1] + Qs[..., 0])
    dpq0, dpq3 = dp[:, 0] + dq[:, 0], dp[:, 3] + dq[:, 3]
    d = dpq0 + dpq3
    on = edge[:, 1:1 + len(xs)] & (d < beta)                    # edge column 0 is the picture border (never filtered)

    def dsam(line, dpq):
        p, q = Ps[:, line], Qs[:, line]
        return (2 * dpq < (beta >> 2)) & (np.abs(p[..., 3] - p[..., 0]) + np.abs(q[..., 0] - q[..., 3]) < (beta >> 3)) & \
               (np.abs(p[..., 0] - q[..., 0]) < ((5 * tc + 1) >> 1))

    strong = on & dsam(0, dpq0) & dsam(3, dpq3)
    weak = on & ~strong
    side = (beta + (beta >> 1)) >> 3
    dep = weak & (dp[:, 0] + dp[:, 3] < side)
    deq = weak & (dq[:, 0] + dq[:, 3] < side)
    rep = lambda m: np.repeat(m, 4, axis=0)                      # per segment -> per line
    strong, weak, dep, deq = rep(strong), rep(weak), rep(dep), rep(deq)
    p0, p1, p2, p3 = (P[..., i] for i in range(4))
    q0, q1, q2, q3 = (Q[..., i] for i in range(4))
    c2 = lambda v, ref: np.clip(v, ref - 2 * tc, ref + 2 * tc)
    sp0 = c2((p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3, p0)
    sp1 = c2((p2 + p1 + p0 + q0 + 2) >> 2, p1)
    sp2 = c2((2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3, p2)
    sq0 = c2((p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3, q0)
    sq1 = c2((p0 + q0 + q1 + q2 + 2) >> 2, q1)
    sq2 = c2((p0 + q0 + q1 + 3 * q2 + 2 * q3 + 4) >> 3, q2)
    delta = (9 * (q0 - p0) - 3 * (q1 - p1) + 8) >> 4
    wk = weak & (np.abs(delta) < tc * 10)
    dl = np.clip(delta, -tc, tc)
    wp0, wq0 = np.clip(p0 + dl, 0, 255), np.clip(q0 - dl, 0, 255)
    tc2 = tc >> 1
    wp1 = np.clip(p1 + np.clip((((p2 + p0 + 1) >> 1) - p1 + dl) >> 1, -tc2, tc2), 0, 255)
    wq1 = np.clip(q1 + np.clip((((q2 + q0 + 1) >> 1) - q1 - dl) >> 1, -tc2, tc2), 0, 255)
    np0 = np.where(strong, sp0, np.where(wk, wp0, p0))
    np1 = np.where(strong, sp1, np.where(wk & dep, wp1, p1))
    np2 = np.where(strong, sp2, p2)
    nq0 = np.where(strong, sq0, np.where(wk, wq0, q0))
    nq1 = np.where(strong, sq1, np.where(wk & deq, wq1, q1))
    nq2 = np.where(strong, sq2, q2)
    out[:, xs - 1], out[:, xs - 2], out[:, xs - 3] = np0, np1, np2
    out[:, xs], out[:, xs + 1], out[:, xs + 2] = nq0, nq1, nq2
    return out


def deblock(rcon, cu_size, kind, qpd6):
    qp = 6 * qpd6 + 4
    beta, tc = BETA[min(max(qp, 0), 51)], TC[min(max(qp + 2, 0), 53)]
    if beta == 0 and tc == 0:
        return rcon.copy()
    tu = tu_size_map(cu_size, kind)                              # (H/4, W/4)
    h4, w4 = tu.shape
    x4 = np.arange(0, w4, 2)
    vedge = ((x4[None, :] * 4) % tu[:, x4]) == 0                 # (H/4 segments, W/8 edges): x is a TU (or PU) boundary
    y4 = np.arange(0, h4, 2)
    hedge = ((y4[:, None] * 4) % tu[y4, :]) == 0                 # (H/8 edges, W/4 segments)
    pic = _filter_edges(rcon.astype(np.int32), vedge, beta, tc)  # all vertical edges first, then horizontal (8.7.2)
    pic = _filter_edges(pic.T.copy(), hedge.T.copy(), beta, tc).T
    return pic.astype(np.uint8)
