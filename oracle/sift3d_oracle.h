/* sift3d_oracle.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * CPU restatement ("port") of the SIFT3D hot path, written from the algorithm
 * description in SURVEY.md Appendix A and checked line by line against the
 * reference sources cited in sift3d_oracle.c.  It is pinned (tests/test_oracle.py)
 * against the unmodified reference compiled into oracle/_ref/ and against the
 * fixtures under tests/golden/ that were generated from that build.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference leg may load this library.
 */
#ifndef SIFT3D_ORACLE_H
#define SIFT3D_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_DESC_NUMEL 768

typedef struct OrcParams {
    double peak_thresh;   /* sift.c:34  default 0.1  */
    double corner_thresh; /* sift.c:36  default 0.4  */
    double sigma_n;       /* sift.c:37  default 1.15 */
    double sigma0;        /* sift.c:38  default 1.6  */
    int num_kp_levels;    /* sift.c:35  default 3    */
} OrcParams;

typedef struct OrcLevel {
    int nx, ny, nz;
    double ux, uy, uz;
    double s;          /* absolute scale of the level */
    const float *data; /* [z][y][x], owned by the context */
} OrcLevel;

typedef struct OrcKeypoint {
    double xd, yd, zd; /* coordinates in the keypoint's own octave */
    double sd;         /* absolute scale */
    int o, s;          /* octave, level */
    float R[9];        /* row-major rotation matrix (columns = principal axes) */
} OrcKeypoint;

typedef struct OrcCtx OrcCtx;

void orc_default_params(OrcParams *p);
OrcCtx *orc_create(const OrcParams *p);
void orc_destroy(OrcCtx *c);

/* SIFT3D_detect_keypoints (sift.c:1609-1641). Returns 0 on success, -1 on failure. */
int orc_detect(OrcCtx *c, const float *vol, int nx, int ny, int nz, const double units[3]);
int orc_num_octaves(const OrcCtx *c);
/* which: 0 = Gaussian pyramid (s = -1..K+1), 1 = DoG pyramid (s = -1..K) */
int orc_get_level(const OrcCtx *c, int which, int o, int s, OrcLevel *out);
int orc_num_candidates(const OrcCtx *c);
const OrcKeypoint *orc_candidates(const OrcCtx *c); /* before orientation rejection */
int orc_num_keypoints(const OrcCtx *c);
const OrcKeypoint *orc_keypoints(const OrcCtx *c);

/* SIFT3D_extract_descriptors (sift.c:2025-2046) on the pyramid left by orc_detect.
 * desc: n x 768 floats; coords (optional): n x 4 doubles (x, y, z in octave-0 voxels, sd). */
int orc_describe(const OrcCtx *c, const OrcKeypoint *kp, int n, float *desc, double *coords);

/* SIFT3D_extract_dense_descriptors, dense_rotate = 0 (sift.c:2354-2496).
 * out: [z][y][x][12] floats. */
int orc_dense(const OrcCtx *c, const float *vol, int nx, int ny, int nz, const double units[3],
              float *out);
/* ... with dense_rotate selectable (rotate = 1: sift.c:2521-2588, :2295-2343) */
int orc_dense_ex(const OrcCtx *c, const float *vol, int nx, int ny, int nz, const double units[3],
                 int rotate, float *out);

/* building blocks, exposed for unit tests */
int orc_gauss_width(double sigma);                      /* imutil.c:3671-3674 */
int orc_gauss_taps(double sigma, float *taps, int cap); /* imutil.c:3657-3710 */
/* apply_Sep_FIR_filter (imutil.c:3459-3544): x, then y, then z; nc interleaved channels */
void orc_blur(const float *src, float *dst, int nx, int ny, int nz, int nc, const double units[3],
              const float *taps, int width, double unit);
float orc_scale(float *data, long n);                     /* im_scale, imutil.c:1977-1991 */
int orc_eig3(const double A[9], double Q[9], double L[3]); /* stand-in for dsyevd, see .c */
/* icosahedron table exactly as init_geometry builds it (sift.c:215-326):
 * v: 20 faces x 3 vertices x 3 coords, idx: 20 x 3 */
void orc_mesh(float *v, int *idx);
/* im_inv_transform for an Affine given as row-major 3x4 (imutil.c:2040-2081, :2651-2672);
 * interp 0 = resample_linear (imutil.c:2085), 1 = resample_lanczos2 (imutil.c:2127) */
int orc_resample_affine(const float *src, int nx, int ny, int nz, int nc, const double A[12],
                        int interp, float *dst, int dnx, int dny, int dnz);
/* SIFT3D_nn_match (sift.c:2840-2969); d1, d2: n x 768 floats */
int orc_nn_match(const float *d1, int n1, const float *d2, int n2, float nn_thresh, int *matches);

#ifdef __cplusplus
}
#endif
#endif
