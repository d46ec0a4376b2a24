/* hevce_variants.h -- private: the kernel variants linked into libhevce_b200.so (see hevce_variant.cu, Makefile). */
#ifndef HEVCE_VARIANTS_H
#define HEVCE_VARIANTS_H

typedef struct hevce_variant_info {
    int gang;                  /* pictures per CTA                                   */
    int threads_per_picture;   /* CTA size = gang * threads_per_picture              */
    int lanes_per_warp;        /* trial-coder lanes hosted by one warp               */
    int wide;                  /* 1: one picture per CTA with the large pool         */
    int tracks;                /* thread tracks per picture (3 = parent || child)    */
    int cluster;               /* CTAs (SMs) per picture: 1, or 2 = tracks on a cluster */
    long long smem_bytes;      /* dynamic shared memory per CTA                      */
} hevce_variant_info;

/* tag, pictures per CTA, threads per picture, lanes per warp, wide pool */
#define HEVCE_VARIANT_LIST(X) \
    X(g7, 7, 128, 32, 0)      \
    X(g4, 4, 224, 24, 0)      \
    X(g2, 2, 448, 16, 0)       \
    X(w1, 1, 896, 8, 1)       \
    X(t1, 1, 896, 12, 0)      \
    X(c2, 1, 896, 8, 0)

#ifdef __cplusplus
extern "C" {
#endif
#define HEVCE_DECLARE_VARIANT(tag, g, nt, lpw, wide)                                                                     \
    int hevce_variant_prepare_##tag(void);                                                                               \
    int hevce_variant_launch_##tag(const void *jobs, const int *gangs, int ngangs, const void *slots, int *counter,      \
                                   const void *tables, int grid, void *stream);                                          \
    void hevce_variant_info_##tag(hevce_variant_info *out);                                                              \
    void hevce_variant_profile_##tag(unsigned long long *cycles, unsigned long long *count);
HEVCE_VARIANT_LIST(HEVCE_DECLARE_VARIANT)
#undef HEVCE_DECLARE_VARIANT
#ifdef __cplusplus
}
#endif
#endif
