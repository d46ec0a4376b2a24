// hevce_variant.cu -- one instantiation of the decision kernel (hevce_core.h: encode_picture).
//
// The library carries several variants of the same kernel source, compiled from this file with different
// -DHEVCE_OPT_GANG / -DHEVCE_OPT_NT / -DHEVCE_OPT_LPW / -DHEVCE_OPT_WIDE (csrc/Makefile): how many pictures share a CTA,
// how many threads each picture gets, how trial-coder lanes are spread over warps and how large the per-picture
// shared-memory pool is.  The session layer (hevce_cuda.cu) picks the variant per batch: many same-size pictures ->
// gangs of 7 in lock-step (throughput, integer-issue bound); few or large pictures -> one picture per CTA with all
// threads and nearly all shared memory of the SM (latency).
//
// Kernel: a persistent grid, one CTA per SM; every CTA pulls gangs of up to GANG same-size pictures from a queue.
// Pictures are independent, so the grid needs no inter-CTA communication.
#include <cuda_runtime.h>

#include "hevce_core.h"
#include "hevce_variants.h"

using namespace HEVCE_NS;

#define HEVCE_CAT2(a, b) a##b
#define HEVCE_CAT(a, b) HEVCE_CAT2(a, b)
#define VSYM(name) HEVCE_CAT(name, HEVCE_VARIANT)

#if defined(HEVCE_PROFILE)
namespace HEVCE_NS {
__device__ unsigned long long g_phase_cycles[128];
__device__ unsigned long long g_phase_count[128];
}
#endif

static constexpr size_t kSmemBytes = NBLOCK * sizeof(Shared) + sizeof(Tables) + sizeof(GangCtl);

// `gangs` lists GANG job indices per work unit, live pictures first, -1 for the empty slots of a short gang.
__global__ void __launch_bounds__(NT * GANG, 1)
VSYM(hevce_encode_kernel_)(const Job* __restrict__ jobs, const int* __restrict__ gangs, int ngangs, const Scratch* __restrict__ slots,
                           int* counter, const Tables* __restrict__ tables) {
    for (int i = threadIdx.x; i < (int)(sizeof(Tables) / 4); i += NT * GANG) ((u32*)&my_tb())[i] = ((const u32*)tables)[i];
#if HEVCE_OPT_CLUSTER
    // one picture per cluster of two CTAs: rank 0 pulls the queue, both CTAs run encode_picture (their own tracks of it)
    const Scratch* scs = slots + (size_t)(blockIdx.x / HEVCE_OPT_CLUSTER) * NTRACK;
    for (;;) {
        if (cluster_rank() == 0 && threadIdx.x == 0) gang_ctl().next = atomicAdd(counter, 1);
        if (threadIdx.x < 3) { blk_sm(0).rdv_seq[threadIdx.x] = 0; blk_sm(1).rdv_seq[threadIdx.x] = 0; }
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
        unsigned long long ctl0;
        asm volatile("mapa.u64 %0, %1, %2;" : "=l"(ctl0) : "l"((unsigned long long)&gang_ctl()), "r"(0));
        const int k = ((const volatile GangCtl*)ctl0)->next;
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
        if (k >= ngangs) break;
        encode_picture(jobs[gangs[k]], scs);
    }
#else
    const int member = threadIdx.x / NT;
    const Scratch* scs = slots + (size_t)(blockIdx.x * GANG + member) * NTRACK;   // this picture slot's scratch, one set per track
    for (;;) {
        if (threadIdx.x == 0) {
            const int k = atomicAdd(counter, 1);
            int live = 0;
            if (k < ngangs)
                for (int m = 0; m < GANG; m++) live += gangs[k * GANG + m] >= 0;
            gang_ctl().next = k;
            gang_ctl().nlive = live;
        }
        __syncthreads();
        const int k = gang_ctl().next, live = gang_ctl().nlive;
        __syncthreads();
        if (k >= ngangs) break;
        // the threads of an empty slot go straight to the queue barrier; the live pictures synchronise among themselves
        if (member < live) encode_picture(jobs[gangs[k * GANG + member]], scs);
    }
#endif
}

extern "C" int VSYM(hevce_variant_prepare_)(void) {
    return (int)cudaFuncSetAttribute(VSYM(hevce_encode_kernel_), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
}

// grid: CTAs (a multiple of the cluster size)
extern "C" int VSYM(hevce_variant_launch_)(const void* jobs, const int* gangs, int ngangs, const void* slots, int* counter,
                                           const void* tables, int grid, void* stream) {
#if HEVCE_OPT_CLUSTER
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(NT * GANG);
    cfg.dynamicSmemBytes = kSmemBytes;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = HEVCE_OPT_CLUSTER;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return (int)cudaLaunchKernelEx(&cfg, VSYM(hevce_encode_kernel_), (const Job*)jobs, gangs, ngangs, (const Scratch*)slots, counter, (const Tables*)tables);
#else
    VSYM(hevce_encode_kernel_)<<<grid, NT * GANG, kSmemBytes, (cudaStream_t)stream>>>((const Job*)jobs, gangs, ngangs, (const Scratch*)slots,
                                                                                      counter, (const Tables*)tables);
    return (int)cudaGetLastError();
#endif
}

extern "C" void VSYM(hevce_variant_info_)(hevce_variant_info* out) {
    out->gang = GANG;
    out->threads_per_picture = NT;
    out->lanes_per_warp = LPW;
    out->wide = WIDE ? 1 : 0;
    out->tracks = NTRACK;
    out->cluster = CLUSTER ? HEVCE_OPT_CLUSTER : 1;
    out->smem_bytes = (long long)kSmemBytes;
}

#if defined(HEVCE_PROFILE)
extern "C" void VSYM(hevce_variant_profile_)(unsigned long long* cycles, unsigned long long* count) {
    cudaMemcpyFromSymbol(cycles, g_phase_cycles, sizeof(g_phase_cycles));
    cudaMemcpyFromSymbol(count, g_phase_count, sizeof(g_phase_count));
}
#endif
