/* csv_io.c -- the text writer behind write_Keypoint_store / write_SIFT3D_Descriptor_store
 * (the step after the path).
 *
 * The reference's write_Mat_rm (imutil.c:1343-1421) prints every element with fprintf("%f")
 * (gzprintf for .csv.gz) and a ',' or '\n' after it: ~150 ns per field, single thread.  Once
 * detect + describe of a 512^3 volume take 80 ms, printing its 60 136 x 771 descriptor matrix
 * that way takes seconds.  Here rows are formatted in parallel blocks (OpenMP, written in
 * order) by a formatter that produces the BYTES printf("%f") produces:
 *
 *   a finite double is M * 2^e with M < 2^53.  M * 10^6 < 2^73 fits an unsigned __int128, so
 *   v * 10^6 = (M * 10^6) / 2^k is an exact integer division; printf rounds the exact decimal
 *   expansion to nearest, ties to even (glibc, default rounding mode) -- the same q, remainder
 *   and tie rule give the same digits.  Values of 2^63 / 10^6 and above, NaN and Inf go through
 *   snprintf.  (tests/test_csv_io.py: 4*10^6 doubles incl. every tie pattern m / 2^k,
 *   denormals, negative zero, against snprintf; whole files against the compiled reference.)
 *
 * A .csv.gz is written as a sequence of gzip members, one per block, compressed in parallel like
 * the text is formatted (RFC 1952 2.2: a gzip file is a series of members; zlib's gzread, zcat
 * and Python's gzip concatenate them).
 */
#include <errno.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <zlib.h>

#include "sift3d_host.h"

#define FIELD_MAX 352 /* "%f" of -DBL_MAX: 1 + 309 + 1 + 6 characters */

static const char DIGITS2[201] =
    "0001020304050607080910111213141516171819202122232425262728293031323334353637383940414243444546474849"
    "5051525354555657585960616263646566676869707172737475767778798081828384858687888990919293949596979899";

/* decimal digits of v, most significant first; returns the count (>= 1) */
static int put_u64(char *dst, uint64_t v)
{
    char tmp[24];
    int n = 0, i;
    while (v >= 100) {
        const unsigned r = (unsigned)(v % 100);
        v /= 100;
        tmp[n++] = DIGITS2[2 * r + 1];
        tmp[n++] = DIGITS2[2 * r];
    }
    if (v >= 10) {
        tmp[n++] = DIGITS2[2 * v + 1];
        tmp[n++] = DIGITS2[2 * v];
    } else {
        tmp[n++] = (char)('0' + v);
    }
    for (i = 0; i < n; i++) dst[i] = tmp[n - 1 - i];
    return n;
}

/* The characters of printf("%f", v); returns their count (no terminator). */
int sift3d_b200_format_f(char *dst, double v)
{
    uint64_t bits, mant, q, ip;
    unsigned fp;
    int ex, n = 0, i;
    memcpy(&bits, &v, sizeof(bits));
    ex = (int)((bits >> 52) & 0x7ff);
    mant = bits & 0xfffffffffffffull;
    if (ex == 0x7ff || ex >= 1023 + 43) /* NaN, Inf, |v| >= 2^43: q would not fit 64 bits */
        return snprintf(dst, FIELD_MAX, "%f", v);
    if (bits >> 63) dst[n++] = '-';
    if (ex == 0) ex = 1; /* denormal: M = mant, e = -1074 */
    else mant |= 1ull << 52;
    {
        /* v = mant * 2^(ex - 1075); k = 1075 - ex >= 10 here */
        const int k = 1075 - ex;
        const unsigned __int128 N = (unsigned __int128)mant * 1000000u;
        if (k >= 128) {
            q = 0; /* N < 2^73: far below half a unit */
        } else {
            const unsigned __int128 one = 1;
            const unsigned __int128 rem = N & ((one << k) - 1), half = one << (k - 1);
            q = (uint64_t)(N >> k);
            if (rem > half || (rem == half && (q & 1))) q++;
        }
    }
    ip = q / 1000000u;
    fp = (unsigned)(q % 1000000u);
    n += put_u64(dst + n, ip);
    dst[n++] = '.';
    for (i = 0; i < 3; i++) {
        const unsigned d = i == 0 ? fp / 10000u : (i == 1 ? fp / 100u % 100u : fp % 100u);
        dst[n++] = DIGITS2[2 * d];
        dst[n++] = DIGITS2[2 * d + 1];
    }
    return n;
}

static int format_d(char *dst, int v)
{
    int n = 0;
    uint64_t u = v < 0 ? (uint64_t)(-(int64_t)v) : (uint64_t)v;
    if (v < 0) dst[n++] = '-';
    return n + put_u64(dst + n, u);
}

/* mkdir -p of the directory part of `path` (mkpath, imutil.c:4145; out_mode 0755, imutil.c:99) */
static int make_parent_dirs(const char *path)
{
    char *copy = strdup(path), *p;
    int rc = 0;
    if (copy == NULL) return -1;
    for (p = copy + 1; *p && !rc; p++) {
        if (*p != '/') continue;
        *p = '\0';
        if (mkdir(copy, 0755) != 0 && errno != EEXIST) rc = -1;
        *p = '/';
    }
    free(copy);
    return rc;
}

typedef struct {
    char *p;
    size_t len, cap;
} Buf;

static int buf_room(Buf *b, size_t extra)
{
    if (b->len + extra <= b->cap) return 0;
    {
        size_t cap = b->cap ? 2 * b->cap : (size_t)1 << 16;
        char *q;
        while (cap < b->len + extra) cap *= 2;
        if ((q = (char *)realloc(b->p, cap)) == NULL) return -1;
        b->p = q;
        b->cap = cap;
    }
    return 0;
}

static int format_rows(const Mat_rm *mat, int r0, int r1, Buf *b)
{
    const int cols = mat->num_cols;
    int i, j;
    b->len = 0;
    for (i = r0; i < r1; i++)
        for (j = 0; j < cols; j++) {
            const size_t q = (size_t)i * cols + j;
            if (buf_room(b, FIELD_MAX + 1)) return -1;
            if (mat->type == SIFT3D_DOUBLE)
                b->len += (size_t)sift3d_b200_format_f(b->p + b->len, mat->u.data_double[q]);
            else if (mat->type == SIFT3D_FLOAT)
                b->len += (size_t)sift3d_b200_format_f(b->p + b->len, (double)mat->u.data_float[q]);
            else
                b->len += (size_t)format_d(b->p + b->len, mat->u.data_int[q]);
            b->p[b->len++] = j < cols - 1 ? ',' : '\n';
        }
    return 0;
}

/* One gzip member holding b (windowBits 15 + 16): concatenated members are a valid .gz stream, so
 * the blocks of a .csv.gz are compressed in parallel like the text is formatted. */
static int gzip_member(const Buf *b, Buf *out)
{
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (deflateInit2(&zs, Z_DEFAULT_COMPRESSION, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK)
        return -1;
    out->len = 0;
    if (buf_room(out, deflateBound(&zs, (uLong)b->len) + 64)) {
        deflateEnd(&zs);
        return -1;
    }
    zs.next_in = (Bytef *)b->p;
    zs.avail_in = (uInt)b->len;
    zs.next_out = (Bytef *)out->p;
    zs.avail_out = (uInt)out->cap;
    if (deflate(&zs, Z_FINISH) != Z_STREAM_END) {
        deflateEnd(&zs);
        return -1;
    }
    out->len = out->cap - zs.avail_out;
    return deflateEnd(&zs) == Z_OK ? 0 : -1;
}

/* write_Mat_rm (imutil.c:1343-1421): .csv, or .csv.gz through zlib. */
int sift3d_b200_write_Mat_rm(const char *path, const Mat_rm *const mat)
{
    const size_t plen = strlen(path);
    const int compress = plen > 3 && strcmp(path + plen - 3, ".gz") == 0;
    FILE *file = NULL;
    int fail = 0, nblocks, rows_per_block, blk;
    if (mat->type != SIFT3D_DOUBLE && mat->type != SIFT3D_FLOAT && mat->type != SIFT3D_INT)
        return SIFT3D_FAILURE;
    if (make_parent_dirs(path)) return SIFT3D_FAILURE;
    if ((file = fopen(path, "w")) == NULL) return SIFT3D_FAILURE;
    rows_per_block = mat->num_cols > 0 ? 32768 / mat->num_cols : 1;
    if (rows_per_block < 1) rows_per_block = 1;
    nblocks = mat->num_cols > 0 ? (mat->num_rows + rows_per_block - 1) / rows_per_block : 0;
    if (compress && nblocks == 0) { /* an empty matrix is still a valid (empty) gzip stream */
        Buf empty = {NULL, 0, 0}, z = {NULL, 0, 0};
        if (gzip_member(&empty, &z) || fwrite(z.p, 1, z.len, file) != z.len) fail = 1;
        free(z.p);
    }
#pragma omp parallel
    {
        Buf b = {NULL, 0, 0}, z = {NULL, 0, 0};
#pragma omp for ordered schedule(static, 1)
        for (blk = 0; blk < nblocks; blk++) {
            const int r0 = blk * rows_per_block;
            const int r1 = r0 + rows_per_block < mat->num_rows ? r0 + rows_per_block : mat->num_rows;
            int bad = format_rows(mat, r0, r1, &b);
            const Buf *w = &b;
            if (!bad && compress) {
                bad = gzip_member(&b, &z);
                w = &z;
            }
#pragma omp ordered
            {
                if (bad) fail = 1;
                else if (!fail && fwrite(w->p, 1, w->len, file) != w->len) fail = 1;
            }
        }
        free(b.p);
        free(z.p);
    }
    if (ferror(file)) fail = 1;
    if (fclose(file) != 0) fail = 1;
    return fail ? SIFT3D_FAILURE : SIFT3D_SUCCESS;
}
