// hevce_simgang.cpp -- TEST INFRASTRUCTURE ONLY: the kernel source (csrc/hevce_core.h) compiled for the host with one
// thread per picture of a gang and real barriers (CTA-wide and per team).  Where hevce_sim.cpp checks one picture's
// logic, this checks what only exists with several pictures in a CTA: trial lanes packed across pictures, the threads a
// lane runs on, work handed to another picture's idle threads, and the barrier sequence of the two teams.  A thread
// stands for the 128 GPU threads of its picture slot and executes exactly the lane / thread indices they own.
#define HEVCE_SIM_GANG 1
#include <pthread.h>

#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "hevce_core.h"

namespace HEVCE_NS {
int g_sim_order = 0;
int g_sim_nlive = 1;
Shared* g_sim_sms = nullptr;
Tables* g_sim_tb = nullptr;
CommitShared* g_sim_csm = nullptr;
thread_local int g_sim_member = 0, g_sim_trk = 0;
static pthread_barrier_t g_bar[4];   // 1 / 2: the teams, 3: the live pictures
void sim_barrier(int id, int) { pthread_barrier_wait(&g_bar[id]); }
}   // namespace HEVCE_NS

// Encode `n` (1..GANG) pictures of identical size h x w as ONE gang (the slots of a short gang stay empty: only the live
// pictures take part in the barriers, as in the kernel).  outs[i] / rcons[i]: per-picture buffers (cap bytes / padded size); lens[i], errs[i] are written.
extern "C" int hevce_simgang_encode(int n, unsigned char* const* outs, int out_cap, const unsigned char* const* imgs,
                                    unsigned char* const* rcons, int h, int w, const int* qs, int order, int* lens, int* errs) {
    using namespace HEVCE_NS;
    if (n < 1 || n > GANG) return -1;
    g_sim_order = order;
    static Tables tables;
    fill_tables(tables);
    g_sim_tb = &tables;
    const int H = (h + CTU - 1) / CTU * CTU, W = (w + CTU - 1) / CTU * CTU, nctu = (H / CTU) * (W / CTU);
    std::vector<Job> jobs(GANG);
    std::vector<std::vector<s16>> glev(GANG), lev(GANG);
    std::vector<std::vector<CtuRec>> recs(GANG);
    std::vector<std::vector<u8>> grec(GANG), line(GANG);
    std::vector<Scratch> sc(GANG);
    std::vector<int> result(2 * GANG, 0);
    for (int m = 0; m < n; m++) {
        Job& j = jobs[m];
        j.img = imgs[m]; j.src_h = h; j.src_w = w; j.H = H; j.W = W; j.q = qs[m]; j.out_cap = out_cap;
        j.out = outs[m]; j.rcon = rcons[m];
        j.result = &result[2 * m];
        glev[m].resize((size_t)NCAND * LEV_STRIDE + 16); lev[m].resize((size_t)nctu * CTU * CTU); recs[m].resize(nctu);
        grec[m].resize((size_t)NREC * CTU * CTU); line[m].resize(W / 4 + 8);
        j.recs = recs[m].data(); j.levs = lev[m].data();
        sc[m].glev = glev[m].data(); sc[m].grec = grec[m].data(); sc[m].msz_line = line[m].data();
    }
    Shared* sms = (Shared*)aligned_alloc(16, sizeof(Shared) * GANG);
    memset(sms, 0xA5, sizeof(Shared) * GANG);
    g_sim_sms = sms;
    g_sim_nlive = n;
    for (int b = 0; b < 4; b++) pthread_barrier_init(&g_bar[b], nullptr, n);
    std::vector<std::thread> th;
    for (int m = 0; m < n; m++)
        th.emplace_back([&, m] {
            g_sim_member = m;
            g_sim_trk = 0;
            encode_picture(jobs[m], &sc[m]);
        });
    for (auto& t : th) t.join();
    for (int b = 0; b < 4; b++) pthread_barrier_destroy(&g_bar[b]);
    free(sms);
    CommitShared* cs = new CommitShared;
    memset(cs, 0x5A, sizeof(CommitShared));
    cs->tb = tables;
    g_sim_csm = cs;
    for (int m = 0; m < n; m++) {
        for (int c = 0; c < nctu; c++) commit_ctu(jobs[m], c, (c * 5 + m) % NTC);
        lens[m] = result[2 * m];
        errs[m] = result[2 * m + 1];
    }
    delete cs;
    return 0;
}

extern "C" int hevce_simgang_size() { return HEVCE_NS::GANG; }
