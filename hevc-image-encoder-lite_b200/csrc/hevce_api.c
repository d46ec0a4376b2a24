/*
 * hevce_api.c -- the C API of libhevce_b200.so: plain C host code over the CUDA layer (hevce_cuda.cu).
 *
 *   HEVCImageEncoder       drop-in for the reference entry point (HEVCe.h:5-12, HEVCe.c:1570-1647)
 *   HEVCImageEncoderBatch  n independent pictures, sharded over the selected GPUs by cumulative CTU count,
 *                          one host thread per device, no collective
 *
 * There is no CPU encoder in this library: every picture is encoded by the sm_100a kernel; if no CUDA device can be
 * used the calls fail with HEVCE_ERR_CUDA.
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "hevce_internal.h"

#define API __attribute__((visibility("default")))
#define MAX_DEV 64
#define CHUNK_PIXELS (768LL * 1024 * 1024)   /* per-device sub-batch bound (padded pixels) so staging stays modest */
#define CHUNK_PICTURES 32768                  /* ... and pictures (the commit kernel's grid.y is the picture index) */

static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;
static int g_ndev = -1, g_devs[MAX_DEV];
static int g_max_dim = 8192;
static hevce_session *g_pool[MAX_DEV];          /* one cached session (grow-only buffers) per device ordinal */
static pthread_mutex_t g_pool_lock[MAX_DEV];
static int g_pool_init = 0;

API const char *hevce_version(void) { return "hevce-b200 0.1 (sm_100a)"; }

int hevce_internal_max_dim(void) { return g_max_dim; }

API int hevce_set_max_dim(int max_dim) {
    int old = g_max_dim;
    if (max_dim >= 32 && max_dim <= 16384) g_max_dim = max_dim;
    return old;
}

static void init_pool_locks(void) {
    int i;
    if (g_pool_init) return;
    for (i = 0; i < MAX_DEV; i++) pthread_mutex_init(&g_pool_lock[i], NULL);
    g_pool_init = 1;
}

static int resolve_devices(void) {   /* call with g_lock held */
    int visible, i;
    init_pool_locks();
    if (g_ndev >= 0) return g_ndev;
    visible = hevce_internal_device_count();
    g_ndev = 0;
    {
        const char *env = getenv("HEVCE_DEVICES");
        if (env && *env) {
            char *copy = strdup(env), *save = NULL, *tok;
            for (tok = strtok_r(copy, ",", &save); tok && g_ndev < MAX_DEV; tok = strtok_r(NULL, ",", &save)) {
                int d = atoi(tok);
                if (d >= 0 && d < visible) g_devs[g_ndev++] = d;
            }
            free(copy);
            return g_ndev;
        }
    }
    for (i = 0; i < visible && i < MAX_DEV; i++) g_devs[g_ndev++] = i;
    return g_ndev;
}

API int hevce_set_devices(int count, const int *ordinals) {
    int i, visible = hevce_internal_device_count();
    if (count < 1 || count > MAX_DEV || !ordinals) return HEVCE_ERR_ARG;
    for (i = 0; i < count; i++)
        if (ordinals[i] < 0 || ordinals[i] >= visible) return visible ? HEVCE_ERR_ARG : HEVCE_ERR_CUDA;
    pthread_mutex_lock(&g_lock);
    init_pool_locks();
    g_ndev = count;
    for (i = 0; i < count; i++) g_devs[i] = ordinals[i];
    pthread_mutex_unlock(&g_lock);
    return 0;
}

typedef struct {
    int device, first, count, status;
    unsigned char *const *pbuffers;
    const unsigned char *const *imgs;
    unsigned char *const *rcons;
    const int *ysz, *xsz, *qpd6;
    int *stream_len;
} Shard;

static long long padded_pixels(int h, int w) {
    long long H = ((h < g_max_dim ? h : g_max_dim) + 31) / 32 * 32, W = ((w < g_max_dim ? w : g_max_dim) + 31) / 32 * 32;
    return H * W;
}

/* encode pictures [first, first+count) on one device, in sub-batches of bounded size */
static void *shard_main(void *arg) {
    Shard *sh = (Shard *)arg;
    int done = 0;
    sh->status = 0;
    pthread_mutex_lock(&g_pool_lock[sh->device]);
    while (done < sh->count && sh->status == 0) {
        int a = sh->first + done, m = 0, rc;
        long long px = 0;
        while (done + m < sh->count && m < CHUNK_PICTURES && (m == 0 || px + padded_pixels(sh->ysz[a + m], sh->xsz[a + m]) <= CHUNK_PIXELS)) {
            px += padded_pixels(sh->ysz[a + m], sh->xsz[a + m]);
            m++;
        }
        if (!g_pool[sh->device]) {
            g_pool[sh->device] = hevce_session_create(sh->device, m, sh->ysz + a, sh->xsz + a, sh->qpd6 + a);
            rc = g_pool[sh->device] ? 0 : HEVCE_ERR_CUDA;
        } else
            rc = hevce_session_configure(g_pool[sh->device], m, sh->ysz + a, sh->xsz + a, sh->qpd6 + a);
        if (!rc) rc = hevce_session_upload(g_pool[sh->device], sh->imgs + a);
        if (!rc) rc = hevce_session_encode(g_pool[sh->device]);
        if (!rc) rc = hevce_session_download(g_pool[sh->device], sh->pbuffers + a, sh->rcons + a, sh->stream_len + a);
        sh->status = rc;
        done += m;
    }
    pthread_mutex_unlock(&g_pool_lock[sh->device]);
    return NULL;
}

API int HEVCImageEncoderBatch(int n, unsigned char *const *pbuffers, const unsigned char *const *imgs,
                              unsigned char *const *img_rcons, int *ysz, int *xsz, const int *qpd6, int *stream_len) {
    Shard shards[MAX_DEV];
    pthread_t threads[MAX_DEV];
    int ndev, devs[MAX_DEV], i, k, nshard = 0, status = 0, *lens = stream_len;
    long long total = 0, acc = 0;
    if (n < 0) return HEVCE_ERR_ARG;
    if (n == 0) return 0;
    if (!pbuffers || !imgs || !img_rcons || !ysz || !xsz || !qpd6) return HEVCE_ERR_ARG;
    for (i = 0; i < n; i++)
        if (!pbuffers[i] || !imgs[i] || !img_rcons[i] || ysz[i] <= 0 || xsz[i] <= 0 || qpd6[i] < 0 || qpd6[i] > 4) return HEVCE_ERR_ARG;
    pthread_mutex_lock(&g_lock);
    ndev = resolve_devices();
    memcpy(devs, g_devs, sizeof(int) * (size_t)(ndev > 0 ? ndev : 0));
    pthread_mutex_unlock(&g_lock);
    if (ndev <= 0) {
        fprintf(stderr, "libhevce_b200: no usable CUDA device (this library has no CPU path)\n");
        return HEVCE_ERR_CUDA;
    }
    if (!lens) lens = (int *)malloc(sizeof(int) * (size_t)n);
    if (!lens) return HEVCE_ERR_ARG;
    for (i = 0; i < n; i++) total += padded_pixels(ysz[i], xsz[i]);
    /* contiguous shards with (nearly) equal padded-pixel = CTU counts */
    if (ndev > n) ndev = n;
    for (k = 0, i = 0; k < ndev; k++) {
        int first = i;
        long long want = total * (k + 1) / ndev;
        while (i < n && (k == ndev - 1 || acc + padded_pixels(ysz[i], xsz[i]) / 2 <= want)) acc += padded_pixels(ysz[i], xsz[i]), i++;
        if (i == first) continue;
        shards[nshard].device = devs[k]; shards[nshard].first = first; shards[nshard].count = i - first;
        shards[nshard].pbuffers = pbuffers; shards[nshard].imgs = imgs; shards[nshard].rcons = img_rcons;
        shards[nshard].ysz = ysz; shards[nshard].xsz = xsz; shards[nshard].qpd6 = qpd6; shards[nshard].stream_len = lens;
        nshard++;
    }
    for (k = 1; k < nshard; k++)
        if (pthread_create(&threads[k], NULL, shard_main, &shards[k])) { shards[k].status = HEVCE_ERR_CUDA; threads[k] = 0; shard_main(&shards[k]); }
    shard_main(&shards[0]);
    for (k = 1; k < nshard; k++)
        if (threads[k]) pthread_join(threads[k], NULL);
    for (k = 0; k < nshard; k++)
        if (shards[k].status && !status) status = shards[k].status;
    if (!status)
        for (i = 0; i < n; i++) {   /* size write-back (HEVCe.c:1643-1644) */
            ysz[i] = ((ysz[i] < g_max_dim ? ysz[i] : g_max_dim) + 31) / 32 * 32;
            xsz[i] = ((xsz[i] < g_max_dim ? xsz[i] : g_max_dim) + 31) / 32 * 32;
        }
    if (lens != stream_len) free(lens);
    return status;
}

API int HEVCImageEncoder(unsigned char *pbuffer, const unsigned char *img, unsigned char *img_rcon, int *ysz, int *xsz, const int qpd6) {
    int len = 0, rc;
    unsigned char *pb[1];
    const unsigned char *im[1];
    unsigned char *rc_[1];
    if (!ysz || !xsz) return HEVCE_ERR_ARG;
    pb[0] = pbuffer; im[0] = img; rc_[0] = img_rcon;
    rc = HEVCImageEncoderBatch(1, pb, im, rc_, ysz, xsz, &qpd6, &len);
    return rc ? rc : len;
}
