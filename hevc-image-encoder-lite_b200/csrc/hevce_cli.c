/*
 * hevce_cli.c -- drop-in command line tool over libhevce_b200.so.
 *
 * Same contract as the reference CLI (/root/reference/src/HEVCeMain.c:138-230): positional arguments
 *     HEVCe <in.pgm> <out.h265> [qpd6] [rcon.pgm]
 * where any lone character '0'..'4' anywhere on the command line is qpd6 (default 3, HEVCeMain.c:150-170), binary P5
 * input with maxval <= 255, the same report on stdout (so scripts that parse it, e.g. HEVCeval.py, keep working) and the
 * reconstruction written with the padded size.  Written from scratch: heap buffers sized from the PGM header instead
 * of three 64 MiB static arrays, bulk fread/fwrite, and the encoder's error return is reported.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/hevce.h"

static unsigned char *read_pgm(const char *path, int *h, int *w) {
    FILE *f = fopen(path, "rb");
    unsigned char *px = NULL;
    int maxval = -1, sep;
    size_t n;
    if (!f) return NULL;
    if (fgetc(f) != 'P' || fgetc(f) != '5' || fscanf(f, "%d %d %d", w, h, &maxval) != 3 || maxval > 255 || *w <= 0 || *h <= 0) goto fail;
    sep = fgetc(f);
    if (sep != ' ' && sep != '\n' && sep != '\r' && sep != '\t') goto fail;
    n = (size_t)(*w) * (size_t)(*h);
    px = (unsigned char *)malloc(n);
    if (!px || fread(px, 1, n, f) != n) goto fail;
    fclose(f);
    return px;
fail:
    free(px);
    fclose(f);
    return NULL;
}

static int write_file(const char *path, const char *header, const unsigned char *data, size_t n) {
    FILE *f = fopen(path, "wb");
    int ok;
    if (!f) return -1;
    ok = (!header || fputs(header, f) >= 0) && fwrite(data, 1, n, f) == n;
    return (fclose(f) == 0 && ok) ? 0 : -1;
}

int main(int argc, char **argv) {
    const char *names[3] = {NULL, NULL, NULL};   /* input, stream, reconstruction */
    int qpd6 = -1, nnames = 0, i, h, w, hp, wp, len, y, x;
    unsigned char *img, *rcon, *stream;
    long long sse = 0;
    double mse;
    char hdr[64];

    for (i = 1; i < argc; i++) {
        if (argv[i][0] >= '0' && argv[i][0] <= '4' && argv[i][1] == '\0') qpd6 = argv[i][0] - '0';
        else if (nnames < 3) names[nnames++] = argv[i];
    }
    if (nnames < 2) {
        printf("Usage:\n    %s  <input-image-file(.pgm)>  <output-file(.hevc/.h265)>  [<qpd6>]  [<output-reconstructed-image-file(.pgm)>]\n\n", argv[0]);
        return -1;
    }
    if (qpd6 < 0) qpd6 = 3;

    printf("arguments:\n");
    printf("  input  image file               = %s\n", names[0]);
    printf("  output stream file              = %s\n", names[1]);
    printf("  Qp%%6                            = %d     (Qp=%d)\n", qpd6, qpd6 * 6 + 4);
    if (names[2]) printf("  output reconstructed image file = %s\n", names[2]);

    img = read_pgm(names[0], &h, &w);
    if (!img) { printf("open %s failed\n", names[0]); return -1; }
    printf("  image size                      = %d x %d\n", w, h);
    printf("compressing...\n");

    hp = ((h < 8192 ? h : 8192) + 31) / 32 * 32;
    wp = ((w < 8192 ? w : 8192) + 31) / 32 * 32;
    rcon = (unsigned char *)malloc((size_t)hp * wp);
    stream = (unsigned char *)malloc(256 + 2 * (size_t)hp * wp);
    if (!rcon || !stream) { printf("out of memory\n"); return -1; }
    {
        int ys = h, xs = w;
        len = HEVCImageEncoder(stream, img, rcon, &ys, &xs, qpd6);
        if (len < 0) { printf("HEVCImageEncoder failed (%d): no usable CUDA device or invalid input\n", len); return -1; }
        hp = ys; wp = xs;
    }
    /* distortion over the overlap of the two sizes, MSE floored at 1e-9 (HEVCeMain.c:116-133) */
    {
        const int hm = h < hp ? h : hp, wm = w < wp ? w : wp;
        for (y = 0; y < hm; y++)
            for (x = 0; x < wm; x++) { const long long d = (long long)img[(size_t)y * w + x] - rcon[(size_t)y * wp + x]; sse += d * d; }
        mse = (double)sse / hm / wm;
        if (mse < 1e-9) mse = 1e-9;
    }
    printf("  padded image size               = %d x %d\n", wp, hp);
    printf("  original   length               = %d Bytes\n", wp * hp);
    printf("  compressed length               = %d Bytes\n", len);
    printf("  compression ratio               = %.5f\n", 1.0 * wp * hp / len);
    printf("  bits per pixel                  = %.5f\n", 8.0 * len / (wp * hp));
    printf("  mean square error (MSE)         = %.7lf\n", mse);
    printf("  peak signal/noise ratio (PSNR)  = %.4lf dB\n", 10.0 * log10(255 * 255 / mse));

    if (write_file(names[1], NULL, stream, (size_t)len)) { printf("write file %s failed\n", names[1]); return -1; }
    if (names[2]) {
        snprintf(hdr, sizeof hdr, "P5\n%d %d\n255\n", wp, hp);
        if (write_file(names[2], hdr, rcon, (size_t)hp * wp)) { printf("write file %s failed\n", names[2]); return -1; }
    }
    free(img); free(rcon); free(stream);
    return 0;
}
