/* sift3d_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C restatement of the SIFT3D hot path (reference bbrister/SIFT3D v1.4.6).
 * Each function cites the reference file:line whose arithmetic it restates.  The
 * arithmetic is deliberately "literal": f32 where the reference is f32, f64 where
 * it is f64, the same operation order, no FMA contraction (build with
 * -ffp-contract=off and without -march=native), so that the pyramid is
 * bit-identical to the compiled reference and the candidate set is identical.
 *
 * Differences in STRUCTURE (not arithmetic): flat arrays instead of Image
 * structs; the separable filter walks each axis in place instead of transposing
 * the volume (im_permute is pure data movement); a Jacobi 3x3 eigen-solver stands
 * in for LAPACK dsyevd (eigenvectors agree to ~1e-15; see tests/test_oracle.py).
 *
 * Pinned by tests/test_oracle.py against oracle/_ref (the unmodified reference
 * compiled by oracle/build_ref.sh) and tests/golden/ fixtures generated from it.
 */
#include "sift3d_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAX(a, b) ((a) > (b) ? (a) : (b))
#define ORC_MIN(a, b) ((a) < (b) ? (a) : (b))

/* sift.c:48-58 internal parameters */
static const double k_max_eig_ratio = 0.90;
static const double k_ori_grad_thresh = 1E-10;
static const double k_bary_eps = FLT_EPSILON * 1E1;
static const double k_ori_sig_fctr = 1.5;
static const double k_ori_rad_fctr = 3.0;
static const double k_desc_sig_fctr = 7.071067812;
static const double k_desc_rad_fctr = 2.0;
static const double k_trunc_thresh = 0.2f * 128.0f / ORC_DESC_NUMEL;
static const double k_gr = 1.6180339887;

typedef struct Level {
    int n[3];
    double u[3];
    double s;
    float *d;
} Level;

struct OrcCtx {
    OrcParams p;
    int noct, nlev_g, nlev_d; /* levels per octave */
    Level *g, *dog;
    float tri_v[20][3][3];
    int tri_idx[20][3];
    OrcKeypoint *cand, *kp;
    int ncand, nkp;
};

/* ---------------------------------------------------------------- filters */

int orc_gauss_width(double sigma)
{ /* imutil.c:3671-3674 */
    const int hw = sigma > 0 ? ORC_MAX((int)ceil(sigma * 3.0), 1) : 1;
    return 2 * hw + 1;
}

int orc_gauss_taps(double sigma, float *taps, int cap)
{ /* init_Gauss_filter, imutil.c:3657-3710 */
    const int width = orc_gauss_width(sigma);
    const int hw = width / 2;
    float acc = 0;
    int i;
    if (width > cap) return -1;
    for (i = 0; i < width; i++) {
        double x = (double)i - hw;
        x /= sigma + DBL_EPSILON;
        taps[i] = (float)exp(-0.5 * x * x);
        acc += taps[i];
    }
    for (i = 0; i < width; i++) taps[i] /= acc;
    return width;
}

/* convolve_sep_gen (imutil.c:2274-2393) along one axis of a [z][y][x][c] array.
 * n = length of the axis, st = element stride of the axis.  Out-of-range reads
 * that are undefined behaviour in the reference (only reachable when the axis
 * is shorter than the filter) are clamped. */
static void conv_line(const float *src, float *dst, int n, long st, const float *taps,
                      int width, float uf)
{
    const int hw = width / 2;
    const float conv_eps = 0.1f;
    const int dim_end = n - 1;
    const int uhw = (int)ceilf(hw * uf);
    const int start = uhw, end = n - 1 - (uhw + 1);
    int i, d;

#define ORC_SAMP(cc)                                                         \
    {                                                                        \
        int lo = (int)(cc);                                                  \
        const float frac = (cc) - (float)lo;                                 \
        int hi = lo + 1;                                                     \
        lo = ORC_MIN(ORC_MAX(lo, 0), dim_end);                               \
        hi = ORC_MIN(ORC_MAX(hi, 0), dim_end);                               \
        acc += tap * ((1.0f - frac) * src[lo * st] + frac * src[hi * st]);   \
    }

    for (i = 0; i < n; i++) {
        float acc = 0.0f;
        if (i >= start && i <= end) {
            float c = (float)i; /* carried across taps, imutil.c:2335-2350 */
            for (d = -hw; d <= hw; d++) {
                const float tap = taps[d + hw];
                const float step = d * uf;
                c -= step;
                ORC_SAMP(c);
                c += step;
            }
        } else {
            for (d = -hw; d <= hw; d++) { /* imutil.c:2365-2387 */
                const float tap = taps[d + hw];
                const float step = d * uf;
                float c = (float)i;
                c -= step;
                if ((int)c < 0)
                    c = -c;
                else if ((int)c >= dim_end)
                    c = 2.0f * dim_end - c - conv_eps;
                ORC_SAMP(c);
            }
        }
        dst[i * st] = acc;
    }
#undef ORC_SAMP
}

static void conv_axis(const float *src, float *dst, const int n[3], int nc, int axis,
                      const float *taps, int width, float uf)
{
    const long st[3] = {nc, (long)nc * n[0], (long)nc * n[0] * n[1]};
    const int a1 = (axis + 1) % 3, a2 = (axis + 2) % 3;
    long j;
    const long nlines = (long)n[a1] * n[a2];
#pragma omp parallel for schedule(static)
    for (j = 0; j < nlines; j++) {
        const long i1 = j % n[a1], i2 = j / n[a1];
        const long base = i1 * st[a1] + i2 * st[a2];
        int c;
        for (c = 0; c < nc; c++)
            conv_line(src + base + c, dst + base + c, n[axis], st[axis], taps, width, uf);
    }
}

void orc_blur(const float *src, float *dst, int nx, int ny, int nz, int nc,
              const double units[3], const float *taps, int width, double unit)
{ /* apply_Sep_FIR_filter, imutil.c:3459-3544: x, y, z in that order; the tap
     spacing in voxels along axis i is (float)(unit / units[i]) (imutil.c:2288) */
    const int n[3] = {nx, ny, nz};
    const size_t tot = (size_t)nx * ny * nz * nc;
    float *tmp = (float *)malloc(tot * sizeof(float));
    float *tmp2 = (float *)malloc(tot * sizeof(float));
    conv_axis(src, tmp, n, nc, 0, taps, width, (float)(unit / units[0]));
    conv_axis(tmp, tmp2, n, nc, 1, taps, width, (float)(unit / units[1]));
    conv_axis(tmp2, dst, n, nc, 2, taps, width, (float)(unit / units[2]));
    free(tmp);
    free(tmp2);
}

float orc_scale(float *data, long n)
{ /* im_max_abs + im_scale, imutil.c:1959-1991 */
    float max = 0.0f;
    long i;
    for (i = 0; i < n; i++) {
        const float a = fabsf(data[i]);
        max = ORC_MAX(max, a);
    }
    if (max == 0.0f) return max;
    for (i = 0; i < n; i++) data[i] /= max;
    return max;
}

/* ---------------------------------------------------------------- geometry */

void orc_mesh(float *vout, int *idxout)
{ /* init_geometry, sift.c:215-326.  NOTE the quirk: the outward test is true for
     all 20 faces, and the swap exchanges the vertex VECTORS v[0]<->v[1] but not
     tri->idx[0]<->idx[1] (sift.c:304-308). */
    const float vert[] = {0, 1, k_gr, 0, -1, k_gr, 0, 1, -k_gr, 0, -1, -k_gr,
                          1, k_gr, 0, -1, k_gr, 0, 1, -k_gr, 0, -1, -k_gr, 0,
                          k_gr, 0, 1, -k_gr, 0, 1, k_gr, 0, -1, -k_gr, 0, -1};
    const int faces[] = {0, 1, 8,  0, 8, 4,  0, 4, 5,  0, 5, 9,   0, 9, 1,  1, 6, 8,  8, 6, 10,
                         8, 10, 4, 4, 10, 2, 4, 2, 5,  5, 2, 11,  5, 11, 9, 9, 11, 7, 9, 7, 1,
                         1, 7, 6,  3, 6, 7,  3, 7, 11, 3, 11, 2,  3, 2, 10, 3, 10, 6};
    int i, j;
    for (i = 0; i < 20; i++) {
        float v[3][3], t1[3], t2[3], nrm[3];
        for (j = 0; j < 3; j++) {
            const int id = faces[3 * i + j];
            float mag;
            idxout[3 * i + j] = id;
            v[j][0] = vert[3 * id];
            v[j][1] = vert[3 * id + 1];
            v[j][2] = vert[3 * id + 2];
            mag = sqrtf(v[j][0] * v[j][0] + v[j][1] * v[j][1] + v[j][2] * v[j][2]);
            /* SIFT3D_CVEC_SCALE(v + j, 1.0f / mag) expands to x * 1.0f / mag (sift.c:295) */
            v[j][0] = v[j][0] * 1.0f / mag;
            v[j][1] = v[j][1] * 1.0f / mag;
            v[j][2] = v[j][2] * 1.0f / mag;
        }
        for (j = 0; j < 3; j++) {
            t1[j] = v[2][j] - v[1][j];
            t2[j] = v[1][j] - v[0][j];
        }
        nrm[0] = t1[1] * t2[2] - t1[2] * t2[1];
        nrm[1] = t1[2] * t2[0] - t1[0] * t2[2];
        nrm[2] = t1[0] * t2[1] - t1[1] * t2[0];
        if (nrm[0] * v[0][0] + nrm[1] * v[0][1] + nrm[2] * v[0][2] < 0) {
            for (j = 0; j < 3; j++) {
                const float t = v[0][j];
                v[0][j] = v[1][j];
                v[1][j] = t;
            }
        }
        memcpy(vout + 9 * i, v, sizeof(v));
    }
}

/* cart2bary (sift.c:335-394): Moller-Trumbore in f32, reference operation order */
static int cart2bary(const float cart[3], const float v[3][3], float bary[3], float *k)
{
    float e1[3], e2[3], t[3], p[3], q[3], det, det_inv;
    int j;
    for (j = 0; j < 3; j++) {
        e1[j] = v[1][j] - v[0][j];
        e2[j] = v[2][j] - v[0][j];
    }
    p[0] = cart[1] * e2[2] - cart[2] * e2[1];
    p[1] = cart[2] * e2[0] - cart[0] * e2[2];
    p[2] = cart[0] * e2[1] - cart[1] * e2[0];
    det = e1[0] * p[0] + e1[1] * p[1] + e1[2] * p[2];
    if (fabsf(det) < k_bary_eps) return -1;
    det_inv = 1.0f / det;
    for (j = 0; j < 3; j++) t[j] = v[0][j] * -1.0f;
    q[0] = t[1] * e1[2] - t[2] * e1[1];
    q[1] = t[2] * e1[0] - t[0] * e1[2];
    q[2] = t[0] * e1[1] - t[1] * e1[0];
    bary[1] = det_inv * (t[0] * p[0] + t[1] * p[1] + t[2] * p[2]);
    bary[2] = det_inv * (cart[0] * q[0] + cart[1] * q[1] + cart[2] * q[2]);
    bary[0] = 1.0f - bary[1] - bary[2];
    *k = (e2[0] * q[0] + e2[1] * q[1] + e2[2] * q[2]) * det_inv;
    return 0;
}

/* icos_hist_bin (sift.c:1646-1683): first face in table order hit by the ray */
static int icos_hist_bin(const OrcCtx *c, const float x[3], float bary[3], int *bin)
{
    int i;
    if (x[0] * x[0] + x[1] * x[1] + x[2] * x[2] < k_bary_eps) return -1;
    for (i = 0; i < 20; i++) {
        float k;
        if (cart2bary(x, c->tri_v[i], bary, &k)) continue;
        if (bary[0] < -k_bary_eps || bary[1] < -k_bary_eps || bary[2] < -k_bary_eps || k < 0)
            continue;
        *bin = i;
        return 0;
    }
    return -1;
}

/* 3x3 symmetric eigen-decomposition, ascending eigenvalues, eigenvectors in the
 * COLUMNS of Q (row-major).  Stand-in for LAPACK dsyevd (imutil.c:2992-3075);
 * cyclic Jacobi in f64. */
int orc_eig3(const double Ain[9], double Q[9], double L[3])
{
    double A[3][3], V[3][3];
    int i, j, sweep, order[3] = {0, 1, 2};
    for (i = 0; i < 3; i++)
        for (j = 0; j < 3; j++) {
            A[i][j] = Ain[3 * i + j];
            V[i][j] = i == j;
        }
    for (sweep = 0; sweep < 64; sweep++) {
        const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
        int p, q;
        if (off == 0.0) break;
        for (p = 0; p < 2; p++)
            for (q = p + 1; q < 3; q++) {
                double theta, t, cs, sn, app, aqq, apq;
                int r;
                if (A[p][q] == 0.0) continue;
                theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                cs = 1.0 / sqrt(t * t + 1.0);
                sn = t * cs;
                app = A[p][p];
                aqq = A[q][q];
                apq = A[p][q];
                A[p][p] = app - t * apq;
                A[q][q] = aqq + t * apq;
                A[p][q] = A[q][p] = 0.0;
                for (r = 0; r < 3; r++) {
                    if (r != p && r != q) {
                        const double arp = A[r][p], arq = A[r][q];
                        A[r][p] = A[p][r] = cs * arp - sn * arq;
                        A[r][q] = A[q][r] = sn * arp + cs * arq;
                    }
                }
                for (r = 0; r < 3; r++) {
                    const double vrp = V[r][p], vrq = V[r][q];
                    V[r][p] = cs * vrp - sn * vrq;
                    V[r][q] = sn * vrp + cs * vrq;
                }
            }
    }
    for (i = 0; i < 3; i++)
        for (j = i + 1; j < 3; j++)
            if (A[order[j]][order[j]] < A[order[i]][order[i]]) {
                const int t = order[i];
                order[i] = order[j];
                order[j] = t;
            }
    for (i = 0; i < 3; i++) {
        L[i] = A[order[i]][order[i]];
        for (j = 0; j < 3; j++) Q[3 * j + i] = V[j][order[i]];
    }
    return 0;
}

/* ---------------------------------------------------------------- context */

void orc_default_params(OrcParams *p)
{ /* sift.c:34-38 */
    p->peak_thresh = 0.1;
    p->corner_thresh = 0.4;
    p->sigma_n = 1.15;
    p->sigma0 = 1.6;
    p->num_kp_levels = 3;
}

OrcCtx *orc_create(const OrcParams *p)
{
    OrcCtx *c = (OrcCtx *)calloc(1, sizeof(OrcCtx));
    if (p)
        c->p = *p;
    else
        orc_default_params(&c->p);
    orc_mesh(&c->tri_v[0][0][0], &c->tri_idx[0][0]);
    return c;
}

static void free_levels(OrcCtx *c)
{
    int i;
    for (i = 0; c->g && i < c->noct * c->nlev_g; i++) free(c->g[i].d);
    for (i = 0; c->dog && i < c->noct * c->nlev_d; i++) free(c->dog[i].d);
    free(c->g);
    free(c->dog);
    c->g = c->dog = NULL;
    c->noct = 0;
}

void orc_destroy(OrcCtx *c)
{
    if (!c) return;
    free_levels(c);
    free(c->cand);
    free(c->kp);
    free(c);
}

#define GLEV(c, o, s) (&(c)->g[(o) * (c)->nlev_g + ((s) + 1)])
#define DLEV(c, o, s) (&(c)->dog[(o) * (c)->nlev_d + ((s) + 1)])

static long lev_size(const Level *l) { return (long)l->n[0] * l->n[1] * l->n[2]; }

static void blur_level(const Level *src, Level *dst, const float *taps, int width)
{ /* build_gpyr calls apply_Sep_FIR_filter with unit = 1.0 (sift.c:1002,1013,1021);
     the output takes the source's dims and units (im_copy_dims, imutil.c:3478) */
    memcpy(dst->n, src->n, sizeof(src->n));
    memcpy(dst->u, src->u, sizeof(src->u));
    orc_blur(src->d, dst->d, src->n[0], src->n[1], src->n[2], 1, src->u, taps, width, 1.0);
}

/* assign_eig_ori (sift.c:1354-1514) + threshold (sift.c:1331-1342).
 * returns 0 = accept, 1 = reject */
static int assign_orientation(const Level *im, const float vc[3], double sigma,
                              double corner_thresh, float R[9], double *conf_out)
{
    const double win_radius = sigma * k_ori_rad_fctr;
    const float uxf = (float)im->u[0], uyf = (float)im->u[1], uzf = (float)im->u[2];
    /* IM_LOOP_SPHERE_START, sift.c:96-119 (rad is a double here) */
    const int x_start = ORC_MAX(floorf(vc[0] - win_radius / uxf), 1);
    const int x_end = ORC_MIN(ceilf(vc[0] + win_radius / uxf), im->n[0] - 2);
    const int y_start = ORC_MAX(floorf(vc[1] - win_radius / uyf), 1);
    const int y_end = ORC_MIN(ceilf(vc[1] + win_radius / uyf), im->n[1] - 2);
    const int z_start = ORC_MAX(floorf(vc[2] - win_radius / uzf), 1);
    const int z_end = ORC_MIN(ceilf(vc[2] + win_radius / uzf), im->n[2] - 2);
    const long ys = im->n[0], zs = (long)im->n[0] * im->n[1];
    double A[9] = {0}, Q[9], L[3], corner_score;
    float vdw[3] = {0.0f, 0.0f, 0.0f}, v[2][3], vr[3];
    int x, y, z, i;

    if (conf_out) *conf_out = 0.0;
    for (z = z_start; z <= z_end; z++)
        for (y = y_start; y <= y_end; y++)
            for (x = x_start; x <= x_end; x++) {
                const float dx = ((float)x - vc[0]) * uxf;
                const float dy = ((float)y - vc[1]) * uyf;
                const float dz = ((float)z - vc[2]) * uzf;
                const float sq_dist = dx * dx + dy * dy + dz * dz;
                const float *p = im->d + x + y * ys + z * zs;
                float vd[3], weight;
                if (sq_dist > win_radius * win_radius) continue;
                weight = expf(-0.5 * sq_dist / (sigma * sigma)); /* sift.c:1401 */
                /* IM_GET_GRAD_ISO, sift.c:150-155 + immacros.h:105-111 */
                vd[0] = 0.5f * (p[1] - p[-1]);
                vd[1] = 0.5f * (p[ys] - p[-ys]);
                vd[2] = 0.5f * (p[zs] - p[-zs]);
                vd[0] *= 1.0f / uxf;
                vd[1] *= 1.0f / uyf;
                vd[2] *= 1.0f / uzf;
                A[0] += (double)vd[0] * vd[0] * weight; /* sift.c:1407-1412 */
                A[1] += (double)vd[0] * vd[1] * weight;
                A[2] += (double)vd[0] * vd[2] * weight;
                A[4] += (double)vd[1] * vd[1] * weight;
                A[5] += (double)vd[1] * vd[2] * weight;
                A[8] += (double)vd[2] * vd[2] * weight;
                vd[0] = vd[0] * weight;
                vd[1] = vd[1] * weight;
                vd[2] = vd[2] * weight;
                vdw[0] = vdw[0] + vd[0];
                vdw[1] = vdw[1] + vd[1];
                vdw[2] = vdw[2] + vd[2];
            }
    A[3] = A[1];
    A[6] = A[2];
    A[7] = A[5];
    if (vdw[0] * vdw[0] + vdw[1] * vdw[1] + vdw[2] * vdw[2] < (float)k_ori_grad_thresh) return 1;
    orc_eig3(A, Q, L);
    for (i = 0; i < 2; i++)
        if (fabs(L[i] / L[i + 1]) > k_max_eig_ratio) return 1;
    corner_score = DBL_MAX;
    for (i = 0; i < 2; i++) { /* sift.c:1448-1480 */
        const int eig_idx = 3 - i - 1;
        double d, cos_ang;
        float sgn;
        vr[0] = (float)Q[0 * 3 + eig_idx];
        vr[1] = (float)Q[1 * 3 + eig_idx];
        vr[2] = (float)Q[2 * 3 + eig_idx];
        d = vdw[0] * vr[0] + vdw[1] * vr[1] + vdw[2] * vr[2];
        cos_ang = d / (sqrtf(vr[0] * vr[0] + vr[1] * vr[1] + vr[2] * vr[2]) *
                       sqrtf(vdw[0] * vdw[0] + vdw[1] * vdw[1] + vdw[2] * vdw[2]));
        corner_score = ORC_MIN(corner_score, fabs(cos_ang));
        sgn = d > 0.0 ? 1.0f : -1.0f;
        vr[0] = vr[0] * sgn;
        vr[1] = vr[1] * sgn;
        vr[2] = vr[2] * sgn;
        R[0 * 3 + i] = vr[0];
        R[1 * 3 + i] = vr[1];
        R[2 * 3 + i] = vr[2];
        memcpy(v[i], vr, sizeof(vr));
    }
    R[0 * 3 + 2] = v[0][1] * v[1][2] - v[0][2] * v[1][1]; /* sift.c:1483-1488 */
    R[1 * 3 + 2] = v[0][2] * v[1][0] - v[0][0] * v[1][2];
    R[2 * 3 + 2] = v[0][0] * v[1][1] - v[0][1] * v[1][0];
    if (conf_out) *conf_out = corner_score;
    return corner_score < corner_thresh ? 1 : 0;
}

int orc_detect(OrcCtx *c, const float *vol, int nx, int ny, int nz, const double units[3])
{
    const int K = c->p.num_kp_levels;
    const int nlev_d = K + 2, nlev_g = K + 3; /* sift.c:945-946 */
    const int mind = ORC_MIN(ORC_MIN(nx, ny), nz);
    const int last_octave = (int)log2((double)mind) - 3; /* sift.c:953-955 */
    float first_taps[64], (*oct_taps)[64];
    int first_w, *oct_w;
    int o, s, i, dims[3] = {nx, ny, nz};
    double u[3] = {units[0], units[1], units[2]};
    Level im;
    long cap = 0;

    if (last_octave < 0) return -1; /* sift.c:958-963 */
    free_levels(c);
    c->noct = last_octave + 1;
    c->nlev_g = nlev_g;
    c->nlev_d = nlev_d;
    c->g = (Level *)calloc((size_t)c->noct * nlev_g, sizeof(Level));
    c->dog = (Level *)calloc((size_t)c->noct * nlev_d, sizeof(Level));
    /* resize_Pyramid (imutil.c:3858-3947) + set_scales_Pyramid (imutil.c:3957-3992) */
    for (o = 0; o < c->noct; o++) {
        for (s = -1; s < nlev_g - 1; s++) {
            Level *l = GLEV(c, o, s);
            memcpy(l->n, dims, sizeof(dims));
            memcpy(l->u, u, sizeof(u));
            l->s = c->p.sigma0 * pow(2.0, o + (double)s / K);
            l->d = (float *)malloc(lev_size(l) * sizeof(float));
        }
        for (s = -1; s < nlev_d - 1; s++) {
            Level *l = DLEV(c, o, s);
            memcpy(l->n, dims, sizeof(dims));
            memcpy(l->u, u, sizeof(u));
            l->s = c->p.sigma0 * pow(2.0, o + (double)s / K);
            l->d = (float *)malloc(lev_size(l) * sizeof(float));
        }
        for (i = 0; i < 3; i++) {
            dims[i] /= 2;
            u[i] *= 2;
        }
    }
    if (GLEV(c, 0, -1)->s < c->p.sigma_n) return -1; /* imutil.c:3975-3981 */

    /* make_gss (imutil.c:3752-3802): filters from octave-0 scales only */
    oct_taps = (float(*)[64])malloc((size_t)(nlev_g - 1) * sizeof(*oct_taps));
    oct_w = (int *)malloc((size_t)(nlev_g - 1) * sizeof(int));
    {
        const double s_first = GLEV(c, 0, -1)->s;
        first_w = orc_gauss_taps(sqrt(s_first * s_first - c->p.sigma_n * c->p.sigma_n),
                                 first_taps, 64);
        for (s = -1; s < nlev_g - 2; s++) {
            const double sc = GLEV(c, 0, s)->s, sn = GLEV(c, 0, s + 1)->s;
            oct_w[s + 1] = orc_gauss_taps(sqrt(sn * sn - sc * sc), oct_taps[s + 1], 64);
        }
    }

    /* set_im_SIFT3D (sift.c:883-913): copy + scale to [-1, 1] */
    im.n[0] = nx;
    im.n[1] = ny;
    im.n[2] = nz;
    memcpy(im.u, units, sizeof(im.u));
    im.d = (float *)malloc(lev_size(&im) * sizeof(float));
    memcpy(im.d, vol, lev_size(&im) * sizeof(float));
    orc_scale(im.d, lev_size(&im));

    /* build_gpyr (sift.c:989-1050) */
    blur_level(&im, GLEV(c, 0, -1), first_taps, first_w);
    free(im.d);
    for (o = 0; o < c->noct; o++) {
        for (s = 0; s <= nlev_g - 2; s++)
            blur_level(GLEV(c, o, s - 1), GLEV(c, o, s), oct_taps[s], oct_w[s]);
        if (o != c->noct - 1) { /* im_downsample_2x, imutil.c:1742-1768 */
            const Level *src = GLEV(c, o, ORC_MAX(nlev_g - 2 - 2, -1));
            Level *dst = GLEV(c, o + 1, -1);
            int x, y, z;
            dst->n[0] = (int)floor((double)src->n[0] / 2.0);
            dst->n[1] = (int)floor((double)src->n[1] / 2.0);
            dst->n[2] = (int)floor((double)src->n[2] / 2.0);
            for (z = 0; z < dst->n[2]; z++)
                for (y = 0; y < dst->n[1]; y++)
                    for (x = 0; x < dst->n[0]; x++)
                        dst->d[x + (long)dst->n[0] * (y + (long)dst->n[1] * z)] =
                            src->d[2 * x + (long)src->n[0] * (2 * y + (long)src->n[1] * 2 * z)];
        }
    }
    free(oct_taps);
    free(oct_w);

    /* build_dog (sift.c:1052-1071) */
    for (o = 0; o < c->noct; o++)
        for (s = -1; s < nlev_d - 1; s++) {
            const Level *a = GLEV(c, o, s), *b = GLEV(c, o, s + 1);
            Level *d = DLEV(c, o, s);
            const long n = lev_size(a);
            long j;
            memcpy(d->n, a->n, sizeof(a->n));
            memcpy(d->u, a->u, sizeof(a->u));
            for (j = 0; j < n; j++) d->d[j] = a->d[j] - b->d[j];
        }

    /* detect_extrema (sift.c:1074-1212): 6 face neighbours + 1 below + 1 above */
    free(c->cand);
    c->cand = NULL;
    c->ncand = 0;
    for (o = 0; o < c->noct; o++)
        for (s = 0; s <= nlev_d - 3; s++) {
            const Level *prev = DLEV(c, o, s - 1), *cur = DLEV(c, o, s), *next = DLEV(c, o, s + 1);
            const long n = lev_size(cur), ys = cur->n[0], zs = (long)cur->n[0] * cur->n[1];
            float dogmax = 0.0f, thr;
            long j;
            int x, y, z;
            for (j = 0; j < n; j++) dogmax = ORC_MAX(dogmax, fabsf(cur->d[j]));
            thr = c->p.peak_thresh * dogmax; /* double product narrowed to float, sift.c:1169 */
            for (z = 1; z <= cur->n[2] - 2; z++)
                for (y = 1; y <= cur->n[1] - 2; y++)
                    for (x = 1; x <= cur->n[0] - 2; x++) {
                        const long q = x + y * ys + z * zs;
                        const float *p = cur->d + q;
                        const float v = *p;
                        OrcKeypoint *k;
                        if (!(v > thr || v < -thr)) continue;
                        if (!((v > prev->d[q] && v > p[1] && v > p[-1] && v > p[ys] &&
                               v > p[-ys] && v > p[-zs] && v > p[zs] && v > next->d[q]) ||
                              (v < prev->d[q] && v < p[1] && v < p[-1] && v < p[ys] &&
                               v < p[-ys] && v < p[-zs] && v < p[zs] && v < next->d[q])))
                            continue;
                        if (c->ncand >= cap) {
                            cap = cap ? 2 * cap : 1024;
                            c->cand = (OrcKeypoint *)realloc(c->cand, cap * sizeof(OrcKeypoint));
                        }
                        k = &c->cand[c->ncand++];
                        memset(k, 0, sizeof(*k));
                        k->o = o;
                        k->s = s;
                        k->sd = cur->s;
                        k->xd = x;
                        k->yd = y;
                        k->zd = z;
                    }
        }

    /* assign_orientations (sift.c:1264-1325): reject + stable compaction */
    free(c->kp);
    c->kp = (OrcKeypoint *)malloc(ORC_MAX(c->ncand, 1) * sizeof(OrcKeypoint));
    {
        char *ok = (char *)malloc(ORC_MAX(c->ncand, 1));
#pragma omp parallel for schedule(dynamic, 16)
        for (i = 0; i < c->ncand; i++) {
            OrcKeypoint *k = &c->cand[i];
            const float vc[3] = {(float)k->xd, (float)k->yd, (float)k->zd};
            ok[i] = !assign_orientation(GLEV(c, k->o, k->s), vc, k_ori_sig_fctr * k->sd,
                                        c->p.corner_thresh, k->R, NULL);
        }
        c->nkp = 0;
        for (i = 0; i < c->ncand; i++)
            if (ok[i]) c->kp[c->nkp++] = c->cand[i];
        free(ok);
    }
    return 0;
}

int orc_num_octaves(const OrcCtx *c) { return c->noct; }
int orc_num_candidates(const OrcCtx *c) { return c->ncand; }
const OrcKeypoint *orc_candidates(const OrcCtx *c) { return c->cand; }
int orc_num_keypoints(const OrcCtx *c) { return c->nkp; }
const OrcKeypoint *orc_keypoints(const OrcCtx *c) { return c->kp; }

int orc_get_level(const OrcCtx *c, int which, int o, int s, OrcLevel *out)
{
    const Level *l;
    if (o < 0 || o >= c->noct || s < -1 || s > (which ? c->nlev_d : c->nlev_g) - 2) return -1;
    l = which ? DLEV(c, o, s) : GLEV(c, o, s);
    out->nx = l->n[0];
    out->ny = l->n[1];
    out->nz = l->n[2];
    out->ux = l->u[0];
    out->uy = l->u[1];
    out->uz = l->u[2];
    out->s = l->s;
    out->data = l->d;
    return 0;
}

/* ---------------------------------------------------------------- descriptors */

static void normalize768(float *h)
{ /* normalize_desc, sift.c:1794-1821 */
    double norm = 0.0;
    float norm_inv;
    int i;
    for (i = 0; i < ORC_DESC_NUMEL; i++) norm += (double)h[i] * h[i];
    norm = sqrt(norm) + DBL_EPSILON;
    norm_inv = 1.0f / norm;
    for (i = 0; i < ORC_DESC_NUMEL; i++) h[i] *= norm_inv;
}

/* extract_descrip (sift.c:1834-1928) + SIFT3D_desc_acc_interp (sift.c:1687-1791) */
static void extract_descrip(const OrcCtx *c, const Level *im, const OrcKeypoint *key, float *h)
{
    const float sigma = key->sd * k_desc_sig_fctr;
    const float win_radius = k_desc_rad_fctr * sigma;
    const float desc_half_width = win_radius / sqrt(2);
    const float desc_width = 2.0f * desc_half_width;
    const float desc_hist_width = desc_width / 4;
    const float desc_bin_fctr = 1.0f / desc_hist_width;
    const float uxf = (float)im->u[0], uyf = (float)im->u[1], uzf = (float)im->u[2];
    const float vc[3] = {(float)key->xd, (float)key->yd, (float)key->zd};
    const int x_start = ORC_MAX(floorf(vc[0] - win_radius / uxf), 1);
    const int x_end = ORC_MIN(ceilf(vc[0] + win_radius / uxf), im->n[0] - 2);
    const int y_start = ORC_MAX(floorf(vc[1] - win_radius / uyf), 1);
    const int y_end = ORC_MIN(ceilf(vc[1] + win_radius / uyf), im->n[1] - 2);
    const int z_start = ORC_MAX(floorf(vc[2] - win_radius / uzf), 1);
    const int z_end = ORC_MIN(ceilf(vc[2] + win_radius / uzf), im->n[2] - 2);
    const long ys = im->n[0], zs = (long)im->n[0] * im->n[1];
    float Rt[9];
    int i, j, x, y, z;

    for (i = 0; i < 3; i++)
        for (j = 0; j < 3; j++) Rt[3 * i + j] = key->R[3 * j + i];
    memset(h, 0, ORC_DESC_NUMEL * sizeof(float));

    for (z = z_start; z <= z_end; z++)
        for (y = y_start; y <= y_end; y++)
            for (x = x_start; x <= x_end; x++) {
                float vim[3], vkp[3], vbins[3], grad[3], grot[3], dv[3], bary[3], mag, weight;
                float sq_dist;
                const float *p = im->d + x + y * ys + z * zs;
                int bin, dx, dy, dz;
                vim[0] = ((float)x - vc[0]) * uxf;
                vim[1] = ((float)y - vc[1]) * uyf;
                vim[2] = ((float)z - vc[2]) * uzf;
                sq_dist = vim[0] * vim[0] + vim[1] * vim[1] + vim[2] * vim[2];
                if (sq_dist > win_radius * win_radius) continue;
                for (i = 0; i < 3; i++) /* SIFT3D_MUL_MAT_RM_CVEC, immacros.h:329-341 */
                    vkp[i] = Rt[3 * i] * vim[0] + Rt[3 * i + 1] * vim[1] + Rt[3 * i + 2] * vim[2];
                for (i = 0; i < 3; i++) vbins[i] = (vkp[i] + desc_half_width) * desc_bin_fctr;
                if (vbins[0] < 0 || vbins[1] < 0 || vbins[2] < 0 || vbins[0] >= 4.0f ||
                    vbins[1] >= 4.0f || vbins[2] >= 4.0f)
                    continue;
                grad[0] = 0.5f * (p[1] - p[-1]);
                grad[1] = 0.5f * (p[ys] - p[-ys]);
                grad[2] = 0.5f * (p[zs] - p[-zs]);
                grad[0] *= 1.0f / uxf;
                grad[1] *= 1.0f / uyf;
                grad[2] *= 1.0f / uzf;
                weight = expf(-0.5f * sq_dist / (sigma * sigma)); /* sift.c:1890 */
                for (i = 0; i < 3; i++) grad[i] = grad[i] * weight;
                for (i = 0; i < 3; i++)
                    grot[i] =
                        Rt[3 * i] * grad[0] + Rt[3 * i + 1] * grad[1] + Rt[3 * i + 2] * grad[2];
                for (i = 0; i < 3; i++) dv[i] = vbins[i] - floorf(vbins[i]);
                if (icos_hist_bin(c, grot, bary, &bin)) continue;
                mag = sqrtf(grot[0] * grot[0] + grot[1] * grot[1] + grot[2] * grot[2]);
                for (dx = 0; dx < 2; dx++)
                    for (dy = 0; dy < 2; dy++)
                        for (dz = 0; dz < 2; dz++) {
                            const int bx = (int)vbins[0] + dx, by = (int)vbins[1] + dy,
                                      bz = (int)vbins[2] + dz;
                            float w, *hist;
                            if (bx < 0 || bx >= 4 || by < 0 || by >= 4 || bz < 0 || bz >= 4)
                                continue;
                            hist = h + 12 * (bx + 4 * by + 16 * bz);
                            w = ((dx == 0) ? (1.0f - dv[0]) : dv[0]) *
                                ((dy == 0) ? (1.0f - dv[1]) : dv[1]) *
                                ((dz == 0) ? (1.0f - dv[2]) : dv[2]);
                            hist[c->tri_idx[bin][0]] += mag * w * bary[0];
                            hist[c->tri_idx[bin][1]] += mag * w * bary[1];
                            hist[c->tri_idx[bin][2]] += mag * w * bary[2];
                        }
            }
    normalize768(h);
    for (i = 0; i < ORC_DESC_NUMEL; i++) h[i] = ORC_MIN(h[i], (float)k_trunc_thresh);
    normalize768(h);
}

int orc_describe(const OrcCtx *c, const OrcKeypoint *kp, int n, float *desc, double *coords)
{
    int i;
    if (n < 1 || !c->g) return -1; /* verify_keys / have_gpyr, sift.c:2057, 2034 */
#pragma omp parallel for schedule(dynamic, 4)
    for (i = 0; i < n; i++) {
        const OrcKeypoint *k = &kp[i];
        extract_descrip(c, GLEV(c, k->o, k->s), k, desc + (size_t)i * ORC_DESC_NUMEL);
        if (coords) {
            const double f = ldexp(1.0, k->o); /* sift.c:1851,1922-1925 */
            coords[4 * i + 0] = k->xd * f;
            coords[4 * i + 1] = k->yd * f;
            coords[4 * i + 2] = k->zd * f;
            coords[4 * i + 3] = k->sd;
        }
    }
    return 0;
}

static void normalize12(float *h)
{ /* normalize_hist, sift.c:2246-2264 */
    double norm = 0.0;
    float norm_inv;
    int i;
    for (i = 0; i < 12; i++) norm += (double)h[i] * h[i];
    norm = sqrt(norm) + DBL_EPSILON;
    norm_inv = 1.0f / norm;
    for (i = 0; i < 12; i++) h[i] *= norm_inv;
}

/* extract_dense_descrip_rotate (sift.c:2295-2343): one 12-bin histogram over a sphere of
 * radius desc_rad_fctr * sigma, gradients rotated by R^T, Gaussian window weight. */
static void dense_descrip_rotate(const OrcCtx *c, const Level *im, const float vc[3],
                                 double sigma, const float R[9], float hist[12])
{
    const float win_radius = k_desc_rad_fctr * sigma;
    const float uxf = (float)im->u[0], uyf = (float)im->u[1], uzf = (float)im->u[2];
    const int x_start = ORC_MAX(floorf(vc[0] - win_radius / uxf), 1);
    const int x_end = ORC_MIN(ceilf(vc[0] + win_radius / uxf), im->n[0] - 2);
    const int y_start = ORC_MAX(floorf(vc[1] - win_radius / uyf), 1);
    const int y_end = ORC_MIN(ceilf(vc[1] + win_radius / uyf), im->n[1] - 2);
    const int z_start = ORC_MAX(floorf(vc[2] - win_radius / uzf), 1);
    const int z_end = ORC_MIN(ceilf(vc[2] + win_radius / uzf), im->n[2] - 2);
    const long ys = im->n[0], zs = (long)im->n[0] * im->n[1];
    float Rt[9];
    int i, j, x, y, z;

    for (i = 0; i < 3; i++)
        for (j = 0; j < 3; j++) Rt[3 * i + j] = R[3 * j + i];
    for (i = 0; i < 12; i++) hist[i] = 0.0f;
    for (z = z_start; z <= z_end; z++)
        for (y = y_start; y <= y_end; y++)
            for (x = x_start; x <= x_end; x++) {
                const float dx = ((float)x - vc[0]) * uxf;
                const float dy = ((float)y - vc[1]) * uyf;
                const float dz = ((float)z - vc[2]) * uzf;
                const float sq_dist = dx * dx + dy * dy + dz * dz;
                const float *p = im->d + x + y * ys + z * zs;
                float grad[3], grot[3], bary[3], mag, weight;
                int bin;
                if (sq_dist > win_radius * win_radius) continue;
                grad[0] = 0.5f * (p[1] - p[-1]);
                grad[1] = 0.5f * (p[ys] - p[-ys]);
                grad[2] = 0.5f * (p[zs] - p[-zs]);
                grad[0] *= 1.0f / uxf;
                grad[1] *= 1.0f / uyf;
                grad[2] *= 1.0f / uzf;
                for (i = 0; i < 3; i++)
                    grot[i] =
                        Rt[3 * i] * grad[0] + Rt[3 * i + 1] * grad[1] + Rt[3 * i + 2] * grad[2];
                if (icos_hist_bin(c, grot, bary, &bin)) continue;
                /* magnitude of the UNROTATED gradient (sift.c:2330) */
                mag = sqrtf(grad[0] * grad[0] + grad[1] * grad[1] + grad[2] * grad[2]);
                weight = expf(-0.5f * sq_dist / (sigma * sigma)); /* sift.c:2333, f64 divide */
                hist[c->tri_idx[bin][0]] += mag * weight * bary[0];
                hist[c->tri_idx[bin][1]] += mag * weight * bary[1];
                hist[c->tri_idx[bin][2]] += mag * weight * bary[2];
            }
}

int orc_dense(const OrcCtx *c, const float *vol, int nx, int ny, int nz, const double units[3],
              float *out)
{
    return orc_dense_ex(c, vol, nx, ny, nz, units, 0, out);
}

int orc_dense_ex(const OrcCtx *c, const float *vol, int nx, int ny, int nz,
                 const double units[3], int rotate, float *out)
{ /* SIFT3D_extract_dense_descriptors (sift.c:2354-2496, rotate: :2521-2588) */
    const long n = (long)nx * ny * nz, ys = nx, zs = (long)nx * ny;
    const float uxf = (float)units[0], uyf = (float)units[1], uzf = (float)units[2];
    const float hist_trunc = k_trunc_thresh * ORC_DESC_NUMEL / 12; /* sift.c:2271 */
    float taps[128], *sm, *tmp;
    int w;
    long j;
    int z;

    /* smooth_scale_raw_input (sift.c:1978-2006) */
    sm = (float *)malloc(n * sizeof(float));
    w = orc_gauss_taps(sqrt(c->p.sigma0 * c->p.sigma0 - c->p.sigma_n * c->p.sigma_n), taps, 128);
    orc_blur(vol, sm, nx, ny, nz, 1, units, taps, w, 1.0);
    orc_scale(sm, n);

    if (rotate) { /* extract_dense_descriptors_rotate (sift.c:2521-2588) */
        const double ori_sigma = c->p.sigma0 * k_ori_sig_fctr;
        const double desc_sigma = c->p.sigma0 * k_desc_sig_fctr / 4;
        Level lv;
        lv.n[0] = nx, lv.n[1] = ny, lv.n[2] = nz;
        lv.u[0] = units[0], lv.u[1] = units[1], lv.u[2] = units[2];
        lv.s = 0.0;
        lv.d = sm;
#pragma omp parallel for schedule(dynamic)
        for (z = 0; z < nz; z++) {
            int x, y;
            for (y = 0; y < ny; y++)
                for (x = 0; x < nx; x++) {
                    const float vc[3] = {(float)x, (float)y, (float)z};
                    static const float Id[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
                    float R[9];
                    const int rej = assign_orientation(&lv, vc, ori_sigma, c->p.corner_thresh, R, NULL);
                    dense_descrip_rotate(c, &lv, vc, desc_sigma, rej ? Id : R,
                                         out + 12 * (x + y * ys + z * zs));
                }
        }
        free(sm);
        goto postproc;
    }

    /* extract_dense_descriptors_no_rotate (sift.c:2429-2496) */
    tmp = (float *)calloc((size_t)n * 12, sizeof(float));
#pragma omp parallel for schedule(static)
    for (z = 1; z <= nz - 2; z++) {
        int x, y;
        for (y = 1; y <= ny - 2; y++)
            for (x = 1; x <= nx - 2; x++) {
                const long q = x + y * ys + z * zs;
                const float *p = sm + q;
                float grad[3], bary[3];
                int bin;
                grad[0] = 0.5f * (p[1] - p[-1]);
                grad[1] = 0.5f * (p[ys] - p[-ys]);
                grad[2] = 0.5f * (p[zs] - p[-zs]);
                grad[0] *= 1.0f / uxf;
                grad[1] *= 1.0f / uyf;
                grad[2] *= 1.0f / uzf;
                if (icos_hist_bin(c, grad, bary, &bin)) continue;
                tmp[12 * q + c->tri_idx[bin][0]] = bary[0];
                tmp[12 * q + c->tri_idx[bin][1]] = bary[1];
                tmp[12 * q + c->tri_idx[bin][2]] = bary[2];
            }
    }
    w = orc_gauss_taps(c->p.sigma0 * k_desc_sig_fctr / 4, taps, 128); /* sift.c:2445-2455 */
    {   /* QUIRK: the 12-channel image inherits the units of the caller's `desc`
           Image (im_copy_dims(desc, &temp), sift.c:2451), which the dense driver
           never sets from `in` (sift.c:2375-2380) -- (1,1,1) after init_im, as in
           cli/denseSift3D.c.  So this blur ignores the input's units. */
        const double desc_units[3] = {1.0, 1.0, 1.0};
        orc_blur(tmp, out, nx, ny, nz, 12, desc_units, taps, w, 1.0);
    }
    free(tmp);
    free(sm);

postproc:
    /* postproc_Hist (sift.c:2267-2292), val = raw input intensity (sift.c:2401) */
#pragma omp parallel for schedule(static)
    for (j = 0; j < n; j++) {
        float *h = out + 12 * j;
        int i;
        normalize12(h);
        for (i = 0; i < 12; i++) h[i] = ORC_MIN(h[i], hist_trunc);
        normalize12(h);
        for (i = 0; i < 12; i++) h[i] *= vol[j];
    }
    return 0;
}

/* ------------------------------------------------------------------ resampling (8f N3)
 * im_inv_transform (imutil.c:2040-2081) for an Affine (apply_Affine_xyz, imutil.c:2651-2672)
 * with resample_linear (imutil.c:2085-2124) or resample_lanczos2 (imutil.c:2127-2178).
 * A: row-major 3x4.  Volumes contiguous, channel-interleaved. */
static double orc_lanczos(double x, double a)
{ /* imutil.c:2181-2185 */
    const double pi_x = M_PI * x;
    return a * sin(pi_x) * sin(pi_x / a) / (pi_x * pi_x);
}

int orc_resample_affine(const float *src, int nx, int ny, int nz, int nc, const double A[12],
                        int interp, float *dst, int dnx, int dny, int dnz)
{
    const size_t sxs = nc, sys = (size_t)nc * nx, szs = (size_t)nc * nx * ny;
    int zi;
    if (interp != 0 && interp != 1) return -1;
#pragma omp parallel for schedule(static)
    for (zi = 0; zi < dnz; zi++) {
        int xi, yi, c;
        for (yi = 0; yi < dny; yi++)
            for (xi = 0; xi < dnx; xi++) {
                const double xd = xi, yd = yi, zd = zi;
                const double x = A[0] * xd + A[1] * yd + A[2] * zd + A[3];
                const double y = A[4] * xd + A[5] * yd + A[6] * zd + A[7];
                const double z = A[8] * xd + A[9] * yd + A[10] * zd + A[11];
                float *out = dst + (size_t)nc * (xi + (size_t)dnx * (yi + (size_t)dny * zi));
                if (x < 0 || x > nx - 1 || y < 0 || y > ny - 1 || z < 0 || z > nz - 1) {
                    for (c = 0; c < nc; c++) out[c] = 0.0f;
                    continue;
                }
                for (c = 0; c < nc; c++) {
                    const float *p = src + c;
                    if (interp == 0) {
                        const int fx = (int)floor(x), fy = (int)floor(y), fz = (int)floor(z);
                        const int cx = (int)ceil(x), cy = (int)ceil(y), cz = (int)ceil(z);
                        const double dist_x = x - fx, dist_y = y - fy, dist_z = z - fz;
                        const double c0 = p[fx * sxs + fy * sys + fz * szs];
                        const double c1 = p[fx * sxs + cy * sys + fz * szs];
                        const double c2 = p[cx * sxs + fy * sys + fz * szs];
                        const double c3 = p[cx * sxs + cy * sys + fz * szs];
                        const double c4 = p[fx * sxs + fy * sys + cz * szs];
                        const double c5 = p[fx * sxs + cy * sys + cz * szs];
                        const double c6 = p[cx * sxs + fy * sys + cz * szs];
                        const double c7 = p[cx * sxs + cy * sys + cz * szs];
                        out[c] = (float)(c0 * (1.0 - dist_x) * (1.0 - dist_y) * (1.0 - dist_z) +
                                         c1 * (1.0 - dist_x) * dist_y * (1.0 - dist_z) +
                                         c2 * dist_x * (1.0 - dist_y) * (1.0 - dist_z) +
                                         c3 * dist_x * dist_y * (1.0 - dist_z) +
                                         c4 * (1.0 - dist_x) * (1.0 - dist_y) * dist_z +
                                         c5 * (1.0 - dist_x) * dist_y * dist_z +
                                         c6 * dist_x * (1.0 - dist_y) * dist_z +
                                         c7 * dist_x * dist_y * dist_z);
                    } else {
                        const double a = 2;
                        const int x0 = (int)ORC_MAX(floor(x) - a, 0.0), x1 = (int)ORC_MIN(floor(x) + a, (double)(nx - 1));
                        const int y0 = (int)ORC_MAX(floor(y) - a, 0.0), y1 = (int)ORC_MIN(floor(y) + a, (double)(ny - 1));
                        const int z0 = (int)ORC_MAX(floor(z) - a, 0.0), z1 = (int)ORC_MIN(floor(z) + a, (double)(nz - 1));
                        double val = 0.0;
                        int xs, ys, zs;
                        for (zs = z0; zs <= z1; zs++)
                            for (ys = y0; ys <= y1; ys++)
                                for (xs = x0; xs <= x1; xs++) {
                                    const double xw = fabs((double)xs - x) + DBL_EPSILON;
                                    const double yw = fabs((double)ys - y) + DBL_EPSILON;
                                    const double zw = fabs((double)zs - z) + DBL_EPSILON;
                                    const double k = orc_lanczos(xw, a) * orc_lanczos(yw, a) * orc_lanczos(zw, a);
                                    val += k * p[xs * sxs + ys * sys + zs * szs];
                                }
                        out[c] = (float)val;
                    }
                }
            }
    }
    return 0;
}

/* ------------------------------------------------------------------ matching (8f N1)
 * SIFT3D_nn_match / match_desc (sift.c:2840-2969).  desc: n x 768 floats, row-major. */
static int orc_match_desc(const float *d, const float *store, int n, float nn_thresh)
{
    double ssd_best = DBL_MAX, ssd_nearest = DBL_MAX;
    int best = -1, i, j;
    for (i = 0; i < n; i++) {
        const float *d2 = store + (size_t)768 * i;
        double ssd = 0.0;
        for (j = 0; j < 768; j++) {
            const double diff = (double)d[j] - (double)d2[j];
            ssd += diff * diff;
            if (j % 12 == 11 && ssd > ssd_nearest) break; /* early exit per histogram */
        }
        if (ssd < ssd_best) {
            best = i;
            ssd_nearest = ssd_best;
            ssd_best = ssd;
        } else {
            ssd_nearest = ORC_MIN(ssd_nearest, ssd);
        }
    }
    if (ssd_best / ssd_nearest > nn_thresh * nn_thresh) return -1;
    return best;
}

int orc_nn_match(const float *d1, int n1, const float *d2, int n2, float nn_thresh, int *matches)
{
    int i;
    if (n1 < 1) return -1;
#pragma omp parallel for schedule(dynamic, 8)
    for (i = 0; i < n1; i++) {
        int m = orc_match_desc(d1 + (size_t)768 * i, d2, n2, nn_thresh);
        if (m >= 0 && orc_match_desc(d2 + (size_t)768 * m, d1, n1, nn_thresh) != i) m = -1;
        matches[i] = m;
    }
    return 0;
}
