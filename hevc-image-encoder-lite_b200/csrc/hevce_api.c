/*
 * hevce_api.c -- the C API of libhevce_b200.so: plain C host code over the CUDA layer (hevce_cuda.cu).
 *
 *   HEVCImageEncoder       drop-in for the reference entry point (HEVCe.h:5-12, HEVCe.c:1570-1647)
 *   HEVCImageEncoderBatch  n independent pictures, sharded over the selected GPUs by cumulative CTU count,
 *                          no collective
 *
 * A shard (the pictures of one device) is cut into chunks; same-size pictures run as rounds of four chunks on four host
 * threads, each with its own session (stream + pinned staging) and its kernel on a quarter of the SMs, so the
 * host<->device copies of one chunk overlap the kernels of the others.
 *
 * There is no CPU encoder in this library: every picture is encoded by the sm_100a kernels; if no CUDA device can be
 * used the calls fail with HEVCE_ERR_CUDA.
 */
#include <errno.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "hevce_internal.h"

#define API __attribute__((visibility("default")))
#define MAX_DEV 64
#define POOL_PER_DEV 4                        /* cached sessions (grow-only buffers) per device ordinal */
#define CHUNK_PIXELS (768LL * 1024 * 1024)    /* per-chunk bound (padded pixels) so the pinned staging stays modest */
#define CHUNK_PICTURES 32768                  /* ... and pictures (the commit kernel's grid.y is the picture index) */
#define WAVE_PICTURES (148 * 7)               /* pictures of one full wave of the 7-picture variant on a B200 */
#define MAX_WORKERS 4                         /* chunk workers per device (= cached sessions per device) */

static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;   /* device list + session pool */
static int g_ndev = -1, g_devs[MAX_DEV];
static int g_max_dim = 8192;
static hevce_session *g_pool[MAX_DEV][POOL_PER_DEV];
static int g_pool_busy[MAX_DEV][POOL_PER_DEV];

API const char *hevce_version(void) { return "hevce-b200 0.2 (sm_100a)"; }

API int hevce_get_max_dim(void) { return __atomic_load_n(&g_max_dim, __ATOMIC_RELAXED); }
int hevce_internal_max_dim(void) { return hevce_get_max_dim(); }

API int hevce_set_max_dim(int max_dim) {
    int old = hevce_get_max_dim();
    if (max_dim >= 32 && max_dim <= 16384) __atomic_store_n(&g_max_dim, max_dim, __ATOMIC_RELAXED);
    return old;
}

/* HEVCE_DEVICES="0,3": every entry must be a visible ordinal; anything else is an error (nothing is cached) */
static int resolve_devices(void) {   /* call with g_lock held; returns the device count or a negative HEVCE_ERR_* */
    int visible, i;
    if (g_ndev >= 0) return g_ndev;
    visible = hevce_internal_device_count();
    {
        const char *env = getenv("HEVCE_DEVICES");
        if (env && *env) {
            int devs[MAX_DEV], nd = 0;
            const char *p = env;
            while (*p) {
                char *end;
                long d;
                errno = 0;
                d = strtol(p, &end, 10);
                if (end == p || errno || d < 0 || d >= visible || nd >= MAX_DEV || (*end && *end != ',')) {
                    fprintf(stderr, "libhevce_b200: HEVCE_DEVICES=\"%s\" is not a list of visible device ordinals (%d visible)\n", env, visible);
                    return visible ? HEVCE_ERR_ARG : HEVCE_ERR_CUDA;
                }
                devs[nd++] = (int)d;
                p = *end ? end + 1 : end;
            }
            if (nd == 0) return HEVCE_ERR_ARG;
            memcpy(g_devs, devs, sizeof(int) * (size_t)nd);
            g_ndev = nd;
            return g_ndev;
        }
    }
    g_ndev = 0;
    for (i = 0; i < visible && i < MAX_DEV; i++) g_devs[g_ndev++] = i;
    return g_ndev;
}

API int hevce_set_devices(int count, const int *ordinals) {
    int i, visible = hevce_internal_device_count();
    if (count < 1 || count > MAX_DEV || !ordinals) return HEVCE_ERR_ARG;
    for (i = 0; i < count; i++)
        if (ordinals[i] < 0 || ordinals[i] >= visible) return visible ? HEVCE_ERR_ARG : HEVCE_ERR_CUDA;
    pthread_mutex_lock(&g_lock);
    g_ndev = count;
    for (i = 0; i < count; i++) g_devs[i] = ordinals[i];
    pthread_mutex_unlock(&g_lock);
    return 0;
}

/* ---- session pool: concurrent callers (and the two chunk workers of a shard) get a session each ---------------- */
static hevce_session *pool_acquire(int device, int *slot) {
    int k;
    hevce_session *s = NULL;
    *slot = -1;
    pthread_mutex_lock(&g_lock);
    for (k = 0; k < POOL_PER_DEV && *slot < 0; k++)
        if (!g_pool_busy[device][k] && g_pool[device][k]) *slot = k;
    for (k = 0; k < POOL_PER_DEV && *slot < 0; k++)
        if (!g_pool_busy[device][k]) *slot = k;
    if (*slot >= 0) { g_pool_busy[device][*slot] = 1; s = g_pool[device][*slot]; }
    pthread_mutex_unlock(&g_lock);
    if (!s) {   /* an empty pool slot, or every pooled session is busy (then the session is temporary) */
        s = hevce_session_create_empty(device);
        if (*slot >= 0) {
            pthread_mutex_lock(&g_lock);
            if (s) g_pool[device][*slot] = s;
            else g_pool_busy[device][*slot] = 0;   /* creation failed: the slot is free again */
            pthread_mutex_unlock(&g_lock);
            if (!s) *slot = -1;
        }
    }
    return s;
}

static void pool_release(int device, int slot, hevce_session *s) {
    if (slot < 0) { hevce_session_destroy(s); return; }
    pthread_mutex_lock(&g_lock);
    g_pool_busy[device][slot] = 0;
    pthread_mutex_unlock(&g_lock);
}

/* Free the cached sessions (HBM buffers, pinned staging) of every device.  Sessions in use by a running call stay. */
API void hevce_release(void) {
    int d, k;
    for (d = 0; d < MAX_DEV; d++)
        for (k = 0; k < POOL_PER_DEV; k++) {
            hevce_session *s = NULL;
            pthread_mutex_lock(&g_lock);
            if (!g_pool_busy[d][k]) { s = g_pool[d][k]; g_pool[d][k] = NULL; }
            pthread_mutex_unlock(&g_lock);
            if (s) hevce_session_destroy(s);
        }
}

typedef struct {
    int device, first, count, status, max_dim, variant, max_ctas;
    unsigned char *const *pbuffers;
    const unsigned char *const *imgs;
    unsigned char *const *rcons;
    const int *ysz, *xsz, *qpd6;
    int *stream_len;
    /* chunk queue shared by the shard's workers */
    int nchunk, next_chunk;
    int *chunk_first;   /* nchunk + 1 entries */
    pthread_mutex_t qlock;
} Shard;

static long long padded_pixels(int h, int w, int max_dim) {
    long long H = ((h < max_dim ? h : max_dim) + 31) / 32 * 32, W = ((w < max_dim ? w : max_dim) + 31) / 32 * 32;
    return H * W;
}

/* worker: encode chunks of the shard until the queue is empty */
static void *chunk_worker(void *arg) {
    Shard *sh = (Shard *)arg;
    int slot;
    hevce_session *s = pool_acquire(sh->device, &slot);
    if (!s) {
        pthread_mutex_lock(&sh->qlock);
        if (!sh->status) sh->status = HEVCE_ERR_CUDA;
        pthread_mutex_unlock(&sh->qlock);
        return NULL;
    }
    for (;;) {
        int c, a, m, rc;
        pthread_mutex_lock(&sh->qlock);
        c = sh->status ? sh->nchunk : sh->next_chunk++;
        pthread_mutex_unlock(&sh->qlock);
        if (c >= sh->nchunk) break;
        a = sh->chunk_first[c];
        m = sh->chunk_first[c + 1] - a;
        rc = hevce_session_configure(s, m, sh->ysz + a, sh->xsz + a, sh->qpd6 + a, sh->max_dim, sh->variant, sh->max_ctas);
        if (!rc) rc = hevce_session_upload(s, sh->imgs + a);
        if (!rc) rc = hevce_session_encode(s);
        if (!rc) rc = hevce_session_download(s, sh->pbuffers + a, sh->rcons + a, sh->stream_len + a);
        if (rc) {
            pthread_mutex_lock(&sh->qlock);
            if (!sh->status) sh->status = rc;
            pthread_mutex_unlock(&sh->qlock);
        }
    }
    pool_release(sh->device, slot, s);
    return NULL;
}

/* encode pictures [first, first+count) on one device */
static void *shard_main(void *arg) {
    Shard *sh = (Shard *)arg;
    int done = 0, cap = 16, nworkers = 2, i, same = 1;
    pthread_t helpers[MAX_WORKERS];
    sh->status = 0;
    sh->nchunk = 0;
    sh->next_chunk = 0;
    sh->chunk_first = (int *)malloc(sizeof(int) * (size_t)(cap + 1));
    if (!sh->chunk_first) { sh->status = HEVCE_ERR_ARG; return NULL; }
    pthread_mutex_init(&sh->qlock, NULL);
    /* the kernel variant is chosen for the whole shard: its chunks share the device */
    sh->variant = hevce_internal_choose_variant(sh->device, sh->count, sh->ysz + sh->first, sh->xsz + sh->first, sh->max_dim);
    for (i = 1; i < sh->count && same; i++)
        if (sh->ysz[sh->first + i] != sh->ysz[sh->first] || sh->xsz[sh->first + i] != sh->xsz[sh->first]) same = 0;
    sh->max_ctas = 0;
    if (same && sh->count >= 8 * 7) {
        /* Same-size pictures: rounds of four chunks (multiples of 7 pictures) on four workers, every chunk's kernel on a
           quarter of the SMs.  The four kernels run side by side, so the copies of a chunk overlap the kernels of the
           others, and the small commit kernel that follows a chunk's decision kernel finds that chunk's SMs free (behind
           whole-device kernels it would wait for the next chunk's persistent CTAs to finish). */
        const long long px = padded_pixels(sh->ysz[sh->first], sh->xsz[sh->first], sh->max_dim) * sh->count;
        int rounds = (int)((px + MAX_WORKERS * CHUNK_PIXELS - 1) / (MAX_WORKERS * CHUNK_PIXELS)), per, nch, sms = hevce_internal_device_sms(sh->device);
        if (rounds < 1) rounds = 1;
        while ((long long)rounds * MAX_WORKERS * (CHUNK_PICTURES - 7) < sh->count) rounds++;   /* many tiny pictures: the per-chunk picture bound */
        nch = rounds * MAX_WORKERS;
        per = (sh->count + nch - 1) / nch;
        per = (per + 6) / 7 * 7;
        nworkers = MAX_WORKERS;
        /* more than one wave: cap every chunk's grid at a quarter of the SMs (smaller shards fit side by side anyway) */
        sh->max_ctas = (sms > 0 && sh->count > WAVE_PICTURES) ? (sms + MAX_WORKERS - 1) / MAX_WORKERS : 0;
        while (nch + 1 > cap) {
            int *p = (int *)realloc(sh->chunk_first, sizeof(int) * (size_t)(2 * cap + 1));
            if (!p) { sh->status = HEVCE_ERR_ARG; break; }
            sh->chunk_first = p;
            cap *= 2;
        }
        while (!sh->status && done < sh->count) {
            sh->chunk_first[sh->nchunk++] = sh->first + done;
            done += per < sh->count - done ? per : sh->count - done;
        }
    }
    while (done < sh->count) {
        int a = sh->first + done, m = 0, same_c = 1;
        long long px = 0;
        while (done + m < sh->count && m < CHUNK_PICTURES &&
               (m == 0 || px + padded_pixels(sh->ysz[a + m], sh->xsz[a + m], sh->max_dim) <= CHUNK_PIXELS)) {
            px += padded_pixels(sh->ysz[a + m], sh->xsz[a + m], sh->max_dim);
            if (sh->ysz[a + m] != sh->ysz[a] || sh->xsz[a + m] != sh->xsz[a]) same_c = 0;
            m++;
        }
        /* same-size pictures: whole waves of the 7-picture variant per chunk, so no chunk ends with a nearly empty wave */
        if (same_c && m > WAVE_PICTURES && done + m < sh->count) m -= m % WAVE_PICTURES;
        if (sh->nchunk == cap) {
            int *p = (int *)realloc(sh->chunk_first, sizeof(int) * (size_t)(2 * cap + 1));
            if (!p) { sh->status = HEVCE_ERR_ARG; break; }
            sh->chunk_first = p;
            cap *= 2;
        }
        sh->chunk_first[sh->nchunk++] = a;
        done += m;
    }
    sh->chunk_first[sh->nchunk] = sh->first + sh->count;
    if (!sh->status) {
        int nh = 0;
        if (nworkers > sh->nchunk) nworkers = sh->nchunk;
        for (i = 1; i < nworkers; i++)
            if (pthread_create(&helpers[nh], NULL, chunk_worker, sh) == 0) nh++;
        chunk_worker(sh);
        for (i = 0; i < nh; i++) pthread_join(helpers[i], NULL);
    }
    pthread_mutex_destroy(&sh->qlock);
    free(sh->chunk_first);
    return NULL;
}

API int HEVCImageEncoderBatch(int n, unsigned char *const *pbuffers, const unsigned char *const *imgs,
                              unsigned char *const *img_rcons, int *ysz, int *xsz, const int *qpd6, int *stream_len) {
    Shard shards[MAX_DEV];
    pthread_t threads[MAX_DEV];
    int ndev, devs[MAX_DEV], started[MAX_DEV], i, k, nshard = 0, status = 0, *lens = stream_len, saved_device;
    const int max_dim = hevce_get_max_dim();   /* one value for the whole call: sharding, clamp and size write-back */
    long long total = 0, acc = 0;
    if (n < 0) return HEVCE_ERR_ARG;
    if (n == 0) return 0;
    if (!pbuffers || !imgs || !img_rcons || !ysz || !xsz || !qpd6) return HEVCE_ERR_ARG;
    for (i = 0; i < n; i++)
        if (!pbuffers[i] || !imgs[i] || !img_rcons[i] || ysz[i] <= 0 || xsz[i] <= 0 || qpd6[i] < 0 || qpd6[i] > 4) return HEVCE_ERR_ARG;
    pthread_mutex_lock(&g_lock);
    ndev = resolve_devices();
    if (ndev > 0) memcpy(devs, g_devs, sizeof(int) * (size_t)ndev);
    pthread_mutex_unlock(&g_lock);
    if (ndev < 0) return ndev;
    if (ndev == 0) {
        fprintf(stderr, "libhevce_b200: no usable CUDA device (this library has no CPU path)\n");
        return HEVCE_ERR_CUDA;
    }
    if (!lens) lens = (int *)malloc(sizeof(int) * (size_t)n);
    if (!lens) return HEVCE_ERR_ARG;
    saved_device = hevce_internal_get_device();   /* shard 0 runs on this thread and selects its device */
    for (i = 0; i < n; i++) total += padded_pixels(ysz[i], xsz[i], max_dim);
    /* contiguous shards with (nearly) equal padded-pixel = CTU counts */
    if (ndev > n) ndev = n;
    for (k = 0, i = 0; k < ndev; k++) {
        int first = i;
        long long want = total * (k + 1) / ndev;
        while (i < n && (k == ndev - 1 || acc + padded_pixels(ysz[i], xsz[i], max_dim) / 2 <= want)) acc += padded_pixels(ysz[i], xsz[i], max_dim), i++;
        if (i == first) continue;
        shards[nshard].device = devs[k]; shards[nshard].first = first; shards[nshard].count = i - first; shards[nshard].max_dim = max_dim;
        shards[nshard].pbuffers = pbuffers; shards[nshard].imgs = imgs; shards[nshard].rcons = img_rcons;
        shards[nshard].ysz = ysz; shards[nshard].xsz = xsz; shards[nshard].qpd6 = qpd6; shards[nshard].stream_len = lens;
        nshard++;
    }
    for (k = 1; k < nshard; k++) {
        started[k] = pthread_create(&threads[k], NULL, shard_main, &shards[k]) == 0;
        if (!started[k]) shard_main(&shards[k]);   /* no thread: run the shard here */
    }
    shard_main(&shards[0]);
    for (k = 1; k < nshard; k++)
        if (started[k]) pthread_join(threads[k], NULL);
    for (k = 0; k < nshard; k++)
        if (shards[k].status && !status) status = shards[k].status;
    if (!status)
        for (i = 0; i < n; i++) {   /* size write-back (HEVCe.c:1643-1644) */
            ysz[i] = ((ysz[i] < max_dim ? ysz[i] : max_dim) + 31) / 32 * 32;
            xsz[i] = ((xsz[i] < max_dim ? xsz[i] : max_dim) + 31) / 32 * 32;
        }
    if (lens != stream_len) free(lens);
    hevce_internal_set_device(saved_device);      /* leave the caller's current device as it was */
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
