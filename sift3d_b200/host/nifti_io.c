/* nifti_io.c -- NIfTI-1 reader / writer for SIFT3D `Image`s.
 *
 * Provides the two symbols the reference's IO dispatcher calls for .nii / .nii.gz
 * (im_read -> read_nii, imutil.c:1230; im_write -> write_nii, imutil.c:1273) with the
 * semantics of imutil/nifti.c:51-221, but on zlib alone -- the reference needs nifticlib,
 * which its own build treats as optional (nifti.c:15-30 are error stubs without it), so a
 * stock kpSift3D in this image cannot even open its example data.  Built into
 * lib/libsift3d_nifti.so; host code by nature (gunzip + a byte-order/scale pass), it sits
 * either side of the accelerated path and adds no compute of its own.
 *
 * What read_nii reproduces (nifti.c:51-163 + nifticlib's header conversion):
 *   dims: the array rank is the last dim[] > 1; rank 4 = 3-D with dim[4] channels, rank > 4
 *   is rejected; units = pixdim[1..3] (0 or non-finite -> 1, as nifti_convert_nhdr2nim does);
 *   every voxel = (float)((double)v * slope + inter) with slope 0 -> 1; file order is
 *   channel-planar [c][z][y][x], Image order is channel-interleaved; either byte order.
 * write_nii (nifti.c:167-221): float32, scl_slope 1 / scl_inter 0, pixdim = units, rank 4
 *   with pixdim[4] = 0 when nc > 1, single-file "n+1" layout with vox_offset 352.
 */
#include "../../include/sift3d_abi.h"

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#define NII_ERR(...) fprintf(stderr, __VA_ARGS__)

#pragma pack(push, 1)
typedef struct nii1_header { /* NIfTI-1, 348 bytes */
    int32_t sizeof_hdr;
    char data_type[10], db_name[18];
    int32_t extents;
    int16_t session_error;
    char regular, dim_info;
    int16_t dim[8];
    float intent_p1, intent_p2, intent_p3;
    int16_t intent_code, datatype, bitpix, slice_start;
    float pixdim[8], vox_offset, scl_slope, scl_inter;
    int16_t slice_end;
    char slice_code, xyzt_units;
    float cal_max, cal_min, slice_duration, toffset;
    int32_t glmax, glmin;
    char descrip[80], aux_file[24];
    int16_t qform_code, sform_code;
    float quatern_b, quatern_c, quatern_d, qoffset_x, qoffset_y, qoffset_z;
    float srow_x[4], srow_y[4], srow_z[4];
    char intent_name[16], magic[4];
} nii1_header;
#pragma pack(pop)

typedef char nii1_header_is_348_bytes[sizeof(nii1_header) == 348 ? 1 : -1];

enum { DT_UINT8 = 2, DT_INT16 = 4, DT_INT32 = 8, DT_FLOAT32 = 16, DT_FLOAT64 = 64, DT_INT8 = 256,
       DT_UINT16 = 512, DT_UINT32 = 768, DT_INT64 = 1024, DT_UINT64 = 1280 };

static void swap_bytes(void *p, size_t size, size_t count)
{
    unsigned char *b = (unsigned char *)p;
    size_t i, j;
    if (size < 2) return;
    for (i = 0; i < count; i++, b += size)
        for (j = 0; j < size / 2; j++) {
            const unsigned char t = b[j];
            b[j] = b[size - 1 - j];
            b[size - 1 - j] = t;
        }
}

static void swap_header(nii1_header *h)
{
    swap_bytes(&h->sizeof_hdr, 4, 1);
    swap_bytes(&h->extents, 4, 1);
    swap_bytes(&h->session_error, 2, 1);
    swap_bytes(h->dim, 2, 8);
    swap_bytes(&h->intent_p1, 4, 3);
    swap_bytes(&h->intent_code, 2, 4);
    swap_bytes(h->pixdim, 4, 8);
    swap_bytes(&h->vox_offset, 4, 3);
    swap_bytes(&h->slice_end, 2, 1);
    swap_bytes(&h->cal_max, 4, 4);
    swap_bytes(&h->glmax, 4, 2);
    swap_bytes(&h->qform_code, 2, 2);
    swap_bytes(&h->quatern_b, 4, 6);
    swap_bytes(h->srow_x, 4, 12);
}

static size_t dtype_size(int dt)
{
    switch (dt) {
    case DT_UINT8: case DT_INT8: return 1;
    case DT_INT16: case DT_UINT16: return 2;
    case DT_INT32: case DT_UINT32: case DT_FLOAT32: return 4;
    case DT_INT64: case DT_UINT64: case DT_FLOAT64: return 8;
    default: return 0;
    }
}

static int gz_read_all(gzFile f, void *dst, size_t n)
{
    unsigned char *p = (unsigned char *)dst;
    while (n) {
        const unsigned chunk = n > (1u << 30) ? (1u << 30) : (unsigned)n;
        const int got = gzread(f, p, chunk);
        if (got <= 0) return -1;
        p += got;
        n -= (size_t)got;
    }
    return 0;
}

int read_nii(const char *path, Image *const im)
{
    nii1_header h;
    gzFile f;
    unsigned char *raw = NULL;
    size_t esize, nvox, skip, size;
    double slope, inter;
    int swapped = 0, rank, i;
    int dims[4] = {1, 1, 1, 1};

    if ((f = gzopen(path, "rb")) == NULL) { /* gzopen also reads plain .nii */
        NII_ERR("read_nii: failure loading file %s\n", path);
        return SIFT3D_FAILURE;
    }
    if (gz_read_all(f, &h, sizeof(h))) goto fail_msg;
    if (h.sizeof_hdr != 348) {
        swap_header(&h);
        swapped = 1;
        if (h.sizeof_hdr != 348) goto fail_msg;
    }
    if (memcmp(h.magic, "n+1", 4) != 0) { /* single-file NIfTI-1 only (.nii / .nii.gz) */
        NII_ERR("read_nii: %s is not a single-file NIfTI-1 image \n", path);
        gzclose(f);
        return SIFT3D_FAILURE;
    }
    if (h.dim[0] < 1 || h.dim[0] > 7) goto fail_msg;
    /* rank = last dimension greater than 1 (nifti.c:66-72) */
    for (rank = h.dim[0]; rank > 0; rank--)
        if (h.dim[rank] > 1) break;
    if (rank > 4) {
        NII_ERR("read_nii: file %s has unsupported dimensionality %d\n", path, rank);
        gzclose(f);
        return SIFT3D_FAILURE;
    }
    for (i = 1; i <= 4 && i <= h.dim[0]; i++) dims[i - 1] = h.dim[i] > 0 ? h.dim[i] : 1;
    if ((esize = dtype_size(h.datatype)) == 0) {
        NII_ERR("read_nii: unsupported datatype %d \n", (int)h.datatype);
        gzclose(f);
        return SIFT3D_FAILURE;
    }
    /* units = pixdim; nifticlib replaces 0 / non-finite entries of the used dims by 1 */
    {
        double u[3];
        for (i = 0; i < 3; i++) {
            const float p = h.pixdim[i + 1];
            u[i] = (i + 1 <= h.dim[0] && (p == 0.0f || !isfinite(p))) ? 1.0 : (double)p;
        }
        im->ux = u[0], im->uy = u[1], im->uz = u[2];
    }
    im->nx = dims[0], im->ny = dims[1], im->nz = dims[2];
    im->nc = rank == 4 ? dims[3] : 1;
    im->xs = (size_t)im->nc; /* im_default_stride, imutil.c:1453 */
    im->ys = im->xs * (size_t)im->nx;
    im->zs = im->ys * (size_t)im->ny;
    nvox = (size_t)im->nx * im->ny * im->nz;
    size = nvox * (size_t)im->nc;
    /* untrusted header: the voxel count must fit the reference's own `int` size arithmetic
     * (im_resize, imutil.c:1533) and vox_offset must be a finite value in [352, 2^31) */
    if ((double)im->nx * im->ny * im->nz * im->nc > 2147483647.0) {
        NII_ERR("read_nii: image in %s is too large (%d x %d x %d x %d)\n", path, im->nx, im->ny,
                im->nz, im->nc);
        gzclose(f);
        return SIFT3D_FAILURE;
    }
    if (!isfinite(h.vox_offset) || h.vox_offset < 352.0f || h.vox_offset > 2147483647.0f) {
        if (!(h.vox_offset == 0.0f)) { /* 0: some writers leave it unset for .nii; data follow at 352 */
            NII_ERR("read_nii: invalid vox_offset %g in %s\n", (double)h.vox_offset, path);
            gzclose(f);
            return SIFT3D_FAILURE;
        }
    }
    if (im->size != size || im->data == NULL) { /* im_resize, imutil.c:1523 */
        float *p = (float *)realloc(im->data, size * sizeof(float));
        if (p == NULL) goto fail_msg;
        im->data = p;
        im->size = size;
    }
    /* voxel data start at vox_offset (>= 352); extensions in between are skipped */
    skip = h.vox_offset >= 352.0f ? (size_t)h.vox_offset - 348 : 4;
    {
        unsigned char junk[4096];
        while (skip) {
            const size_t n = skip > sizeof(junk) ? sizeof(junk) : skip;
            if (gz_read_all(f, junk, n)) goto fail_msg;
            skip -= n;
        }
    }
    if ((raw = (unsigned char *)malloc(size * esize)) == NULL) goto fail_msg;
    if (gz_read_all(f, raw, size * esize)) goto fail_msg;
    gzclose(f);
    f = NULL;
    if (swapped) swap_bytes(raw, esize, size);

    slope = (double)h.scl_slope; /* ignore a zero slope: ill-formatted image (nifti.c:99-101) */
    if (slope == 0.0) slope = 1.0;
    inter = (double)h.scl_inter;
#define COPY_FROM(type)                                                                   \
    {                                                                                     \
        const type *src = (const type *)raw;                                              \
        long long c_;                                                                     \
        for (c_ = 0; c_ < im->nc; c_++) {                                                 \
            const type *plane = src + (size_t)c_ * nvox;                                  \
            long long v_;                                                                 \
            _Pragma("omp parallel for schedule(static)")                                  \
            for (v_ = 0; v_ < (long long)nvox; v_++)                                      \
                im->data[(size_t)v_ * im->nc + c_] =                                      \
                    (float)((double)plane[v_] * slope + inter);                           \
        }                                                                                 \
    }
    switch (h.datatype) {
    case DT_UINT8: COPY_FROM(uint8_t) break;
    case DT_INT8: COPY_FROM(int8_t) break;
    case DT_UINT16: COPY_FROM(uint16_t) break;
    case DT_INT16: COPY_FROM(int16_t) break;
    case DT_UINT32: COPY_FROM(uint32_t) break;
    case DT_INT32: COPY_FROM(int32_t) break;
    case DT_UINT64: COPY_FROM(uint64_t) break;
    case DT_INT64: COPY_FROM(int64_t) break;
    case DT_FLOAT32: COPY_FROM(float) break;
    case DT_FLOAT64: COPY_FROM(double) break;
    }
#undef COPY_FROM
    free(raw);
    return SIFT3D_SUCCESS;

fail_msg:
    NII_ERR("read_nii: failure loading file %s\n", path);
    if (f) gzclose(f);
    free(raw);
    return SIFT3D_FAILURE;
}

int write_nii(const char *path, const Image *const im)
{
    nii1_header h;
    const size_t len = strlen(path);
    const int gz = len >= 3 && strcmp(path + len - 3, ".gz") == 0;
    const size_t nvox = (size_t)im->nx * im->ny * im->nz;
    const int multi = im->nc > 1;
    float *planar = NULL;
    int x, y, z, c, ok = 0;

    if (im->data == NULL || im->nx < 1 || im->ny < 1 || im->nz < 1 || im->nc < 1 ||
        im->nx > 32767 || im->ny > 32767 || im->nz > 32767 || im->nc > 32767)
        return SIFT3D_FAILURE;
    memset(&h, 0, sizeof(h));
    h.sizeof_hdr = 348;
    h.regular = 'r';
    h.dim[0] = multi ? 4 : 3;
    h.dim[1] = (int16_t)im->nx, h.dim[2] = (int16_t)im->ny, h.dim[3] = (int16_t)im->nz;
    h.dim[4] = multi ? (int16_t)im->nc : 1;
    h.dim[5] = h.dim[6] = h.dim[7] = 1;
    h.datatype = DT_FLOAT32;
    h.bitpix = 32;
    h.pixdim[0] = 1.0f; /* qfac */
    h.pixdim[1] = (float)im->ux, h.pixdim[2] = (float)im->uy, h.pixdim[3] = (float)im->uz;
    h.pixdim[4] = multi ? 0.0f : 1.0f; /* channels have no size (nifti.c:190-191) */
    h.pixdim[5] = h.pixdim[6] = h.pixdim[7] = 1.0f;
    h.vox_offset = 352.0f;
    h.scl_slope = 1.0f;
    h.scl_inter = 0.0f;
    memcpy(h.magic, "n+1", 4);

    /* Image is channel-interleaved with strides; the file is channel-planar */
    if ((planar = (float *)malloc(nvox * im->nc * sizeof(float))) == NULL) return SIFT3D_FAILURE;
    for (c = 0; c < im->nc; c++)
        for (z = 0; z < im->nz; z++)
            for (y = 0; y < im->ny; y++)
                for (x = 0; x < im->nx; x++)
                    planar[x + (size_t)im->nx * (y + (size_t)im->ny * (z + (size_t)im->nz * c))] =
                        im->data[c + x * im->xs + y * im->ys + z * im->zs];
    {
        const unsigned char extender[4] = {0, 0, 0, 0};
        const size_t nb = nvox * im->nc * sizeof(float);
        if (gz) {
            gzFile f = gzopen(path, "wb1"); /* level 1: float data barely compresses */
            if (f) {
                const unsigned char *p = (const unsigned char *)planar;
                size_t left = nb;
                ok = gzwrite(f, &h, sizeof(h)) == (int)sizeof(h) && gzwrite(f, extender, 4) == 4;
                while (ok && left) {
                    const unsigned chunk = left > (1u << 30) ? (1u << 30) : (unsigned)left;
                    ok = gzwrite(f, p, chunk) == (int)chunk;
                    p += chunk;
                    left -= chunk;
                }
                ok = (gzclose(f) == Z_OK) && ok;
            }
        } else {
            FILE *f = fopen(path, "wb");
            if (f) {
                ok = fwrite(&h, sizeof(h), 1, f) == 1 && fwrite(extender, 4, 1, f) == 1 &&
                     fwrite(planar, 1, nb, f) == nb;
                ok = (fclose(f) == 0) && ok;
            }
        }
    }
    free(planar);
    if (!ok) NII_ERR("write_nii: failed to write %s \n", path);
    return ok ? SIFT3D_SUCCESS : SIFT3D_FAILURE;
}
