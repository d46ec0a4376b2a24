// hevce_simstage.cpp -- TEST INFRASTRUCTURE ONLY: single stages of the kernel source (csrc/hevce_core.h) on the host, so
// that each can be compared with the function of the reference that it replaces (tests/test_stages.py calls the
// reference's exported getBorder / predict / transform / quantize / deQuantize / putCoef + CABAClen through ctypes):
//   hevce_stage_pixel    : reference samples -> prediction -> residual -> forward transform -> RDOQ -> group zero-out ->
//                          dequantisation -> inverse transform -> reconstruction + SSE of ONE candidate (phases border, A-D)
//   hevce_stage_residual : residual_coding() of one TU from a fresh coder and fresh contexts (put_residual)
//   hevce_stage_rdoq     : the per-coefficient RDOQ decision (rdoq_level) for a list of coefficients
//   hevce_stage_tables   : the coder tables of the kernel (bin table, rate steps, initial contexts) as plain integers
#include <cstdlib>
#include <cstring>
#include <vector>

#include "hevce_core.h"

namespace HEVCE_NS {
int g_sim_order = 0, g_sim_nlive = 1;
Shared* g_sim_sms = nullptr;
Tables* g_sim_tb = nullptr;
CommitShared* g_sim_csm = nullptr;
thread_local int g_sim_member = 0, g_sim_trk = 0;
void sim_barrier(int, int) {}
}   // namespace HEVCE_NS

using namespace HEVCE_NS;

static Tables* tables() {
    static Tables t;
    static bool done = false;
    if (!done) { fill_tables(t); done = true; }
    return &t;
}

template <int T>
static void pixel(Shared& sm, const Scratch& sc, int mode, int q, int ty, int tx, const Avail& av) {
    Grp g;
    g.n = 1; g.cand0 = 0; g.mode0 = mode; g.ty = ty; g.tx = tx; g.av = av; g.priv = 0;
    g.cuy = ty; g.cux = tx; g.cus = T; g.tu = 0; g.one_tu = 1; g.grec = 1;
    g.blk = 0; g.pred = 4096; g.psum = 6144; g.bord = 12288; g.rec = -1; g.rec_stride = 0; g.rec_pitch = 0;
    const RdK rk = rd_consts(q);
    for (int j = 0; j <= 4 * T; j++) border_column<T>(sm, sm, g, j, 0, 0);
    for (int i = 0; i < T; i++) phase_a_item<T>(sm, sm, g, i);
    for (int i = 0; i < T; i++) phase_b_item<T>(sm, sm, g, i, q, rk);
    for (int i = 0; i < T; i++) phase_c_item<T>(sm, sc, g, i, q);
    for (int i = 0; i < T; i++) phase_d_item<T>(sm, sm, sc, g, i);
}

// win: the (CTU+1) x WP reconstruction window (row 0 / column 0 = neighbours outside the CTU), orig: the CTU's 32x32
// original; the TU of size T sits at (ty, tx) inside the CTU.  Outputs: pred / rec T*T raster, lev T*T raster (int).
extern "C" int hevce_stage_pixel(int T, int mode, int q, const unsigned char* win, const unsigned char* orig, int ty, int tx,
                                 int aL, int aLB, int aA, int aAR, unsigned char* pred, int* lev, unsigned char* rec, int* sse) {
    Shared* sm = new Shared[NTRACK];
    memset(sm, 0x5c, sizeof(Shared) * NTRACK);
    g_sim_sms = sm; g_sim_tb = tables(); g_sim_trk = 0; g_sim_member = 0;
    memcpy(sm->win, win, sizeof(sm->win));
    memcpy(sm->orig, orig, sizeof(sm->orig));
    std::vector<s16> glev((size_t)NCAND * LEV_STRIDE + 16);
    std::vector<u8> grec((size_t)NREC * CTU * CTU);
    Scratch sc;
    sc.glev = glev.data(); sc.grec = grec.data(); sc.msz_line = nullptr;
    const Avail av = {aL, aLB, aA, aAR};
    if (T == 4) pixel<4>(*sm, sc, mode, q, ty, tx, av);
    else if (T == 8) pixel<8>(*sm, sc, mode, q, ty, tx, av);
    else if (T == 16) pixel<16>(*sm, sc, mode, q, ty, tx, av);
    else pixel<32>(*sm, sc, mode, q, ty, tx, av);
    memcpy(pred, sm->pool + 4096, (size_t)T * T);
    memcpy(rec, grec.data(), (size_t)T * T);
    // levels: groups in raster order, the 16 levels of a group in the scan order of the mode -> raster
    const int st = scan_type(T, mode), ncg = T / 4;
    for (int gy = 0; gy < ncg; gy++)
        for (int gx = 0; gx < ncg; gx++)
            for (int k = 0; k < 16; k++) {
                const int p4 = tables()->scan4[st][k];
                lev[(gy * 4 + (p4 >> 2)) * T + gx * 4 + (p4 & 3)] = glev[(gy * ncg + gx) * 16 + k];
            }
    *sse = sm->cand_sse[0];
    delete[] sm;
    return 0;
}

// The level RDOQ picks for each of n coefficients of a TU of size T at qpd6 = q (before the coefficient-group zero-out).
extern "C" void hevce_stage_rdoq(int T, int q, int n, const int* cf, int* lev) {
    const Tables& tb = *tables();
    const RdoqK k = rdoq_consts(ilog2(T), q);
    for (int i = 0; i < n; i++) {
        int dl;
        const int pick = rdoq_level(cf[i], k, tb, dl);
        lev[i] = cf[i] < 0 ? -pick : pick;
    }
}

// Kernel tables as integers.  st: [128][6] = LPS range for range quarter 0..3, context after an LPS, context after an MPS
// (Tables::st8); drate: [8] (Tables::drate); ctx: [5][4 * CTXW] initial context bytes for qpd6 = 0..4 in the compact layout.
extern "C" int hevce_stage_tables(int* st, int* drate, int* ctx) {
    const Tables& tb = *tables();
    for (int v = 0; v < 128; v++) {
        for (int q = 0; q < 4; q++) st[v * 6 + q] = (int)((tb.st8[v] >> (8 * q)) & 0xff);
        st[v * 6 + 4] = (int)((tb.st8[v] >> 32) & 0xff);
        st[v * 6 + 5] = (int)((tb.st8[v] >> 40) & 0xff);
    }
    for (int i = 0; i < 8; i++) drate[i] = tb.drate[i];
    for (int q = 0; q < 5; q++)
        for (int i = 0; i < 4 * CTXW; i++) ctx[q * 4 * CTXW + i] = ctx_init_value(tb.ctx_iv[i], q);
    return 4 * CTXW;
}

// lev: T*T raster levels.  state: {range, low, nbits, nbytes, held, z, n} at the end.  returns the bits written.
extern "C" int hevce_stage_residual(int T, int mode, int q, const int* lev, int* state) {
    const Tables& tb = *tables();
    g_sim_tb = tables();
    u8 ctx[4 * CTXW];
    for (int i = 0; i < 4 * CTXW; i++) ctx[i] = ctx_init_value(tb.ctx_iv[i], q);
    const int st = scan_type(T, mode), ncg = T / 4;
    std::vector<s16> blocked((size_t)T * T + 16);
    unsigned mlo = 0, mhi = 0;
    for (int gy = 0; gy < ncg; gy++)
        for (int gx = 0; gx < ncg; gx++)
            for (int k = 0; k < 16; k++) {
                const int p4 = tb.scan4[st][k];
                const int v = lev[(gy * 4 + (p4 >> 2)) * T + gx * 4 + (p4 & 3)];
                blocked[(gy * ncg + gx) * 16 + k] = (s16)v;
                if (v) { const int b = gy * 8 + gx; if (b < 32) mlo |= 1u << b; else mhi |= 1u << (b - 32); }
            }
    Bac b;
    coder_reset(b.c);
    b.out = nullptr; b.cap = 0; b.tabs = 0;
    const int len0 = coder_len(b.c);
    const Cx cx = {ctx};
    put_residual(b, tb, cx, T, mode, blocked.data(), mlo, mhi);
    state[0] = b.c.range; state[1] = b.c.low; state[2] = b.c.nbits; state[3] = b.c.nbytes; state[4] = b.c.held; state[5] = b.c.z; state[6] = b.c.n;
    return coder_len(b.c) - len0;
}

// Bypass grouping: the same random sequence of context bins and bypass strings through the byte-writing coder, once with
// every bypass string in ONE put_bypass call and once cut into random pieces (one call each, down to one bit per call as
// the reference does for the last-position suffixes).  Returns 0 when both byte streams (incl. emulation prevention and
// the final flush) are identical, else the 1-based position of the first difference; *len receives the stream length.
extern "C" int hevce_stage_bypass_grouping(unsigned seed, int nitems, int* len) {
    const Tables& tb = *tables();
    auto rnd = [&seed]() { seed = seed * 1664525u + 1013904223u; return seed >> 8; };
    struct Item { int ctx, bin, bits, n; };
    std::vector<Item> items((size_t)nitems);
    for (auto& it : items) {
        if (rnd() % 3 == 0) { it.ctx = (int)(rnd() % NCTX); it.bin = (int)(rnd() & 1); it.n = 0; it.bits = 0; }
        else { it.ctx = -1; it.n = 1 + (int)(rnd() % 30); it.bits = (int)(rnd() ^ (rnd() << 12)) & ((1 << it.n) - 1); it.bin = 0; }
        if (rnd() % 17 == 0 && it.ctx < 0) it.bits = 0;                                  // zero runs: emulation prevention
    }
    std::vector<u8> out[2];
    for (int pass = 0; pass < 2; pass++) {
        u8 ctx[4 * CTXW];
        for (int i = 0; i < 4 * CTXW; i++) ctx[i] = ctx_init_value(tb.ctx_iv[i], 2);
        out[pass].assign((size_t)nitems * 8 + 64, 0);
        BacCommit b;
        b.tabs = 0;
        coder_reset(b.c);
        b.out = out[pass].data(); b.cap = (int)out[pass].size();
        unsigned s2 = seed ^ 0x9e3779b9u;
        auto rnd2 = [&s2]() { s2 = s2 * 1664525u + 1013904223u; return s2 >> 8; };
        for (const auto& it : items) {
            if (it.ctx >= 0) { b.put_bin(tb, it.bin, ctx[it.ctx]); continue; }
            if (pass == 0) { b.put_bypass(it.bits, it.n); continue; }
            for (int left = it.n; left > 0;) {                                           // MSB-first pieces
                const int k = 1 + (int)(rnd2() % (unsigned)imin(left, 5));
                b.put_bypass((it.bits >> (left - k)) & ((1 << k) - 1), k);
                left -= k;
            }
        }
        b.put_terminate(1);
        b.finish();
        out[pass].resize((size_t)b.c.n);
    }
    *len = (int)out[0].size();
    if (out[0].size() != out[1].size()) return 1 + (int)imin((int)out[0].size(), (int)out[1].size());
    for (size_t i = 0; i < out[0].size(); i++) if (out[0][i] != out[1][i]) return 1 + (int)i;
    return 0;
}
