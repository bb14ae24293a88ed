/* sift3d_abi.h -- binary interface of the drop-in libsift3D.so built by this repo.
 *
 * The B200 build replaces the reference's libsift3D.so underneath unchanged
 * callers (kpSift3D, denseSift3D, regSift3D, libreg, MEX files), which
 * stack-allocate these objects.  Every struct below therefore has exactly the
 * size, field order and field meaning of its namesake in the reference
 * (imutil/imtypes.h:136-334, x86-64 SysV); the _Static_asserts at the bottom pin
 * the numbers measured from the compiled reference.
 * Only the types the hot path touches are declared; the function list is the
 * complete sift3d/sift.h:19-108 surface.
 */
#ifndef SIFT3D_ABI_H
#define SIFT3D_ABI_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { SIFT3D_SUCCESS = 0, SIFT3D_FAILURE = -1, SIFT3D_SINGULAR = 1 };
enum { SIFT3D_FALSE = 0, SIFT3D_TRUE = 1 };

#define IM_NDIMS 3
#define NHIST_PER_DIM 4
#define ICOS_NFACES 20
#define ICOS_NVERT 12
#define HIST_NUMEL ICOS_NVERT
#define DESC_NUM_TOTAL_HIST (NHIST_PER_DIM * NHIST_PER_DIM * NHIST_PER_DIM)
#define DESC_NUMEL (DESC_NUM_TOTAL_HIST * HIST_NUMEL)

typedef enum _Mat_rm_type { SIFT3D_DOUBLE, SIFT3D_FLOAT, SIFT3D_INT } Mat_rm_type;

/* dense row-major matrix: 32 bytes */
typedef struct _Mat_rm {
    union {
        double *data_double;
        float *data_float;
        int *data_int;
    } u;
    size_t size; /* bytes */
    int num_cols;
    int num_rows;
    int static_mem; /* data aliases caller memory; never realloc/free */
    Mat_rm_type type;
} Mat_rm;

/* volume, x fastest, channel interleaved: 104 bytes */
typedef struct _Image {
    float *data;
    int cl_image; /* reference: dead OpenCL handle */
    double s;     /* scale-space coordinate */
    size_t size;  /* elements */
    int nx, ny, nz;
    double ux, uy, uz;
    size_t xs, ys, zs; /* strides in elements */
    int nc;
    int cl_valid;
} Image;

typedef struct _Sep_FIR_filter {
    int cl_apply_unrolled;
    float *kernel;
    int dim;
    int width;
    int symmetric;
} Sep_FIR_filter;

typedef struct _Gauss_filter {
    double sigma;
    Sep_FIR_filter f;
} Gauss_filter;

typedef struct _GSS_filters {
    Gauss_filter first_gauss;
    Gauss_filter *gauss_octave;
    int num_filters;
    int first_level;
} GSS_filters;

typedef struct _SIFT_cl_kernels {
    int downsample_2; /* B200 build: engine handle id (see host/sift3d_api.c) */
} SIFT_cl_kernels;

/* scale-space pyramid: 48 bytes; level (o,s) = levels[(o-first_octave)*num_levels + s-first_level] */
typedef struct _Pyramid {
    Image *levels;
    double sigma_n;
    double sigma0;
    int num_kp_levels;
    int first_octave;
    int num_octaves;
    int first_level;
    int num_levels;
} Pyramid;

typedef struct _Cvec {
    float x, y, z;
} Cvec;

typedef struct _Slab {
    void *buf;
    size_t num;
    size_t buf_size; /* bytes */
} Slab;

/* 112 bytes; R.u.data_float must alias r_data (static_mem = 1) */
typedef struct _Keypoint {
    float r_data[IM_NDIMS * IM_NDIMS];
    Mat_rm R;
    double xd, yd, zd;
    double sd;
    int o, s;
} Keypoint;

typedef struct _Keypoint_store {
    Keypoint *buf;
    Slab slab;
    int nx, ny, nz;
} Keypoint_store;

typedef struct _Hist {
    float bins[HIST_NUMEL];
} Hist;

typedef struct _Tri {
    Cvec v[3];
    int idx[3];
} Tri;

typedef struct _Mesh {
    Tri *tri;
    int num;
} Mesh;

/* 3104 bytes */
typedef struct _SIFT3D_Descriptor {
    Hist hists[DESC_NUM_TOTAL_HIST];
    double xd, yd, zd, sd;
} SIFT3D_Descriptor;

typedef struct _SIFT3D_Descriptor_store {
    SIFT3D_Descriptor *buf;
    size_t num;
    int nx, ny, nz;
} SIFT3D_Descriptor_store;

/* 304 bytes */
typedef struct _SIFT3D {
    Mesh mesh;
    GSS_filters gss;
    SIFT_cl_kernels kernels;
    Pyramid gpyr;
    Pyramid dog;
    Image im;
    double peak_thresh;
    double corner_thresh;
    int dense_rotate;
} SIFT3D;

#ifndef __cplusplus
_Static_assert(sizeof(Mat_rm) == 32, "Mat_rm ABI");
_Static_assert(sizeof(Image) == 104 && offsetof(Image, nx) == 32 && offsetof(Image, ux) == 48 &&
                   offsetof(Image, xs) == 72 && offsetof(Image, nc) == 96,
               "Image ABI");
_Static_assert(sizeof(Pyramid) == 48, "Pyramid ABI");
_Static_assert(sizeof(Keypoint) == 112 && offsetof(Keypoint, R) == 40 &&
                   offsetof(Keypoint, xd) == 72 && offsetof(Keypoint, o) == 104,
               "Keypoint ABI");
_Static_assert(sizeof(Keypoint_store) == 48, "Keypoint_store ABI");
_Static_assert(sizeof(SIFT3D_Descriptor) == 3104 && offsetof(SIFT3D_Descriptor, xd) == 3072,
               "SIFT3D_Descriptor ABI");
_Static_assert(sizeof(SIFT3D_Descriptor_store) == 32, "SIFT3D_Descriptor_store ABI");
_Static_assert(sizeof(SIFT3D) == 304 && offsetof(SIFT3D, gss) == 16 &&
                   offsetof(SIFT3D, kernels) == 72 && offsetof(SIFT3D, gpyr) == 80 &&
                   offsetof(SIFT3D, dog) == 128 && offsetof(SIFT3D, im) == 176 &&
                   offsetof(SIFT3D, peak_thresh) == 280 && offsetof(SIFT3D, dense_rotate) == 296,
               "SIFT3D ABI");
#endif

/* ---- sift3d/sift.h:19-108, same names, arguments and return conventions ---- */
void init_Keypoint_store(Keypoint_store *const kp);
int init_Keypoint(Keypoint *const key);
int resize_Keypoint_store(Keypoint_store *const kp, const size_t num);
int copy_Keypoint(const Keypoint *const src, Keypoint *const dst);
void cleanup_Keypoint_store(Keypoint_store *const kp);
void init_SIFT3D_Descriptor_store(SIFT3D_Descriptor_store *const desc);
void cleanup_SIFT3D_Descriptor_store(SIFT3D_Descriptor_store *const desc);
int set_peak_thresh_SIFT3D(SIFT3D *const sift3d, const double peak_thresh);
int set_corner_thresh_SIFT3D(SIFT3D *const sift3d, const double corner_thresh);
int set_num_kp_levels_SIFT3D(SIFT3D *const sift3d, const unsigned int num_kp_levels);
int set_sigma_n_SIFT3D(SIFT3D *const sift3d, const double sigma_n);
int set_sigma0_SIFT3D(SIFT3D *const sift3d, const double sigma0);
int init_SIFT3D(SIFT3D *sift3d);
int copy_SIFT3D(const SIFT3D *const src, SIFT3D *const dst);
void cleanup_SIFT3D(SIFT3D *const sift3d);
void print_opts_SIFT3D(void);
int parse_args_SIFT3D(SIFT3D *const sift3d, const int argc, char **argv, const int check_err);
int SIFT3D_assign_orientations(const SIFT3D *const sift3d, const Image *const im,
                               Keypoint_store *const kp, double **const conf);
int SIFT3D_detect_keypoints(SIFT3D *const sift3d, const Image *const im, Keypoint_store *const kp);
int SIFT3D_have_gpyr(const SIFT3D *const sift3d);
int SIFT3D_extract_descriptors(SIFT3D *const sift3d, const Keypoint_store *const kp,
                               SIFT3D_Descriptor_store *const desc);
int SIFT3D_extract_raw_descriptors(SIFT3D *const sift3d, const Image *const im,
                                   const Keypoint_store *const kp,
                                   SIFT3D_Descriptor_store *const desc);
int SIFT3D_extract_dense_descriptors(SIFT3D *const sift3d, const Image *const in,
                                     Image *const desc);
int SIFT3D_nn_match(const SIFT3D_Descriptor_store *const d1,
                    const SIFT3D_Descriptor_store *const d2, const float nn_thresh,
                    int **const matches);
int Keypoint_store_to_Mat_rm(const Keypoint_store *const kp, Mat_rm *const mat);
int SIFT3D_Descriptor_coords_to_Mat_rm(const SIFT3D_Descriptor_store *const store,
                                       Mat_rm *const mat);
int SIFT3D_Descriptor_store_to_Mat_rm(const SIFT3D_Descriptor_store *const store,
                                      Mat_rm *const mat);
int Mat_rm_to_SIFT3D_Descriptor_store(const Mat_rm *const mat,
                                      SIFT3D_Descriptor_store *const store);
int SIFT3D_matches_to_Mat_rm(SIFT3D_Descriptor_store *d1, SIFT3D_Descriptor_store *d2,
                             const int *const matches, Mat_rm *const match1,
                             Mat_rm *const match2);
int draw_matches(const Image *const left, const Image *const right, const Mat_rm *const keys_left,
                 const Mat_rm *const keys_right, const Mat_rm *const match_left,
                 const Mat_rm *const match_right, Image *const concat, Image *const keys,
                 Image *const lines);
int write_Keypoint_store(const char *path, const Keypoint_store *const kp);
int write_SIFT3D_Descriptor_store(const char *path, const SIFT3D_Descriptor_store *const desc);

/* ---- B200 extensions (not in the reference) -------------------------------- */
/* Materialise pyramid level (o,s) from HBM into host memory (which: 0 Gaussian, 1 DoG). */
int sift3d_b200_fetch_level(const SIFT3D *sift3d, int which, int o, int s, float *dst);
/* Host copies of every level in gpyr.levels[i].data / dog.levels[i].data (malloc memory owned by
 * the SIFT3D object), for callers that read the pyramids as they would after the reference's
 * detect (write_pyramid imutil.c:4093, copy_SIFT3D sift.c:650-651).  $SIFT3D_HOST_PYRAMID=1
 * makes every SIFT3D_detect_keypoints end with it. */
int sift3d_b200_materialize_pyramids(SIFT3D *sift3d);
/* Opaque engine behind a SIFT3D object (s3d_engine*, include/sift3d_cuda.h), or NULL. */
void *sift3d_b200_engine(const SIFT3D *sift3d);
/* Number of candidates found by the last detect (before orientation rejection). */
int sift3d_b200_num_candidates(const SIFT3D *sift3d);
/* SIFT3D_detect_keypoints (sift.c:1609) for ONE volume Z-slab tiled over several GPUs
 * (BASELINE.json configs[4]): `im` holds this rank's planes [zsplit[r], zsplit[r+1]), `comm` is
 * an s3d_comm* (include/sift3d_cuda.h).  Collective; keypoints come back in global coordinates,
 * in the reference's scan order within the rank.  INTEGRATION.md section 5. */
int SIFT3D_detect_keypoints_slab(SIFT3D *const sift3d, const Image *const im, const int *zsplit,
                                 void *comm, Keypoint_store *const kp);
/* im_inv_transform with an Affine (imutil.c:2040-2083; A = the 3 x 4 matrix, row-major; interp:
 * 0 = LINEAR, 1 = LANCZOS2 as in interp_type, imtypes.h:105-108) and im_resample
 * (imutil.c:2191-2244) on the device.  SURVEY.md 8f N3. */
int sift3d_b200_im_inv_transform_affine(const double A[12], const Image *const src, const int interp,
                                        const int resize, Image *const dst);
int sift3d_b200_im_resample(const Image *const src, const double *const units, const int interp,
                            Image *const dst);
/* write_Mat_rm (imutil.c:1343-1421), .csv or .csv.gz: the bytes the reference writes, rows
 * formatted in parallel (host/csv_io.c; write_Keypoint_store / write_SIFT3D_Descriptor_store sit
 * on it).  sift3d_b200_format_f: the characters of printf("%f", v) -- at most 352, no
 * terminator -- and their count.  SURVEY.md 8f N2. */
int sift3d_b200_write_Mat_rm(const char *path, const Mat_rm *const mat);
int sift3d_b200_format_f(char *dst, double v);
/* Host-side Gaussian tap design (init_Gauss_filter imutil.c:3657-3710,
 * init_Gauss_incremental_filter imutil.c:3713-3734): f64 design, f32 normalisation.  The caller
 * frees g->f.kernel. */
int s3dh_gauss_filter(Gauss_filter *g, double sigma, int dim);
int s3dh_gauss_incremental(Gauss_filter *g, double s_cur, double s_next, int dim);

#ifdef __cplusplus
}
#endif
#endif
