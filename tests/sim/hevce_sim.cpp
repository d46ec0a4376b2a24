// hevce_sim.cpp -- TEST INFRASTRUCTURE ONLY: compiles the kernel source (csrc/hevce_core.h) for the host and runs
// one CTA single-threaded, every PAR_FOR phase in a permuted item order.  It exists so the device logic can be
// checked bit-for-bit against the oracle inside the build container (which has no GPU) and so that phases with
// hidden intra-phase dependencies show up as mismatches under permutation.  The product library never links it.
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
}

static std::vector<HEVCE_NS::CtuRec> g_last_recs;
static int g_last_h = 0, g_last_w = 0;

extern "C" int hevce_sim_encode(unsigned char* out, int out_cap, const unsigned char* img, unsigned char* rcon,
                                int* ysz, int* xsz, int q, int order, int max_dim, int* err) {
    using namespace HEVCE_NS;
    g_sim_order = order;
    static Tables tables;
    fill_tables(tables);
    Job job;
    int result[2] = {0, 0};
    job.img = img; job.rcon = rcon; job.out = out; job.result = result;
    job.src_h = *ysz; job.src_w = *xsz;
    job.H = (imin(*ysz, max_dim) + CTU - 1) / CTU * CTU;
    job.W = (imin(*xsz, max_dim) + CTU - 1) / CTU * CTU;
    job.q = q; job.out_cap = out_cap;
    const int nctu = (job.H / CTU) * (job.W / CTU);
    std::vector<s16> lev((size_t)nctu * CTU * CTU);
    std::vector<CtuRec> recs(nctu);
    job.recs = recs.data(); job.levs = lev.data();
    std::vector<std::vector<s16>> glev(NTRACK);
    std::vector<std::vector<u8>> grec(NTRACK), line(NTRACK);
    Scratch sc[NTRACK];
    for (int t = 0; t < NTRACK; t++) {   // one scratch set per track (the tracks run one after the other here)
        glev[t].resize((size_t)NCAND * LEV_STRIDE + 16); grec[t].resize((size_t)NREC * CTU * CTU); line[t].resize(job.W / 4 + 8);
        sc[t].glev = glev[t].data(); sc[t].grec = grec[t].data(); sc[t].msz_line = line[t].data();
    }
    Shared* sm = new Shared[NTRACK];
    memset(sm, 0xA5, sizeof(Shared) * NTRACK);   // shared memory is not zeroed on the GPU either
    g_sim_sms = sm;
    g_sim_tb = &tables;
    g_sim_trk = 0;
    encode_picture(job, sc);
    delete[] sm;
    // commit pass (hevce_commit_kernel on the GPU): one CTU at a time here
    CommitShared* cs = new CommitShared;
    memset(cs, 0x5A, sizeof(CommitShared));
    cs->tb = tables;
    g_sim_csm = cs;
    for (int c = 0; c < nctu; c++) commit_ctu(job, order == 1 ? nctu - 1 - c : c, (c * 7) % NTC);
    delete cs;
    *ysz = job.H; *xsz = job.W;
    g_last_recs = recs; g_last_h = job.H; g_last_w = job.W;
    if (err) *err = result[1];
    return result[0];
}

// decisions of the last hevce_sim_encode call: (H/4)x(W/4) CU sizes and modes, (H/8)x(W/8) CU kinds
extern "C" void hevce_sim_last_partition(unsigned char* cu_size, unsigned char* mode, unsigned char* kind) {
    HEVCE_NS::unpack_partition(g_last_recs.data(), g_last_h, g_last_w, cu_size, mode, kind);
}

extern "C" int hevce_sim_shared_bytes() { return (int)sizeof(HEVCE_NS::Shared); }
