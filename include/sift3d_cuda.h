/* sift3d_cuda.h -- C ABI of the B200 device engine (libsift3d_cuda.so).
 *
 * This is the thin shim the host library (libsift3D.so, plain C) calls.  Every
 * entry point takes plain pointers and sizes; no C++/torch types cross it.  Each
 * function names the reference routine whose work it replaces (file:line in
 * bbrister/SIFT3D v1.4.6) -- the reference has no FFI of its own for this path
 * (single process, host memory only), so "what its FFI would bind" is exactly
 * the set of internal calls SIFT3D_detect_keypoints / SIFT3D_extract_descriptors /
 * SIFT3D_extract_dense_descriptors make into imutil.c and sift.c.
 *
 * Conventions: every function returns 0 on success and -1 on failure (the
 * reference's SIFT3D_SUCCESS / SIFT3D_FAILURE, imtypes.h:20-22); the message is
 * available from s3d_engine_error().  Volumes are [z][y][x] float32, x fastest
 * (SIFT3D_IM_GET_IDX, immacros.h:58-59).  There is NO CPU fallback: if no CUDA
 * device is usable, s3d_engine_create fails.
 */
#ifndef SIFT3D_CUDA_H
#define SIFT3D_CUDA_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define S3D_DESC_NUMEL 768   /* 4^3 spatial cells x 12 icosahedron vertices (imtypes.h:79-98) */
#define S3D_DESC_STRIDE 3104 /* sizeof(SIFT3D_Descriptor): 768 f32 + xd,yd,zd,sd f64 */
#define S3D_MAX_TAPS 129     /* widest separable filter accepted (half width 64) */

typedef struct s3d_engine s3d_engine;

/* Geometry of one pyramid level (the metadata half of `Image`, imtypes.h:156-168). */
typedef struct s3d_geom {
    int nx, ny, nz;
    double ux, uy, uz; /* physical units of a voxel */
    double scale;      /* absolute scale s of the level (Image.s) */
} s3d_geom;

/* One separable FIR filter (Sep_FIR_filter, imtypes.h:171-179); taps on the host. */
typedef struct s3d_filter {
    const float *taps;
    int width;
} s3d_filter;

/* Keypoint record exchanged with the host (the data half of `Keypoint`,
 * imtypes.h:253-261).  x,y,z are the f32 centre the reference uses
 * (`Cvec vcenter = {key->xd, ...}`, sift.c:1280 and sift.c:1866-1868). */
typedef struct s3d_keypoint {
    float R[9]; /* row-major */
    float x, y, z;
    double sd;
    int o, s;
} s3d_keypoint; /* 64 bytes */

/* ---- lifecycle ------------------------------------------------------------ */
/* device < 0: use $SIFT3D_CUDA_DEVICE, else $LOCAL_RANK, else 0. */
int s3d_engine_create(s3d_engine **out, int device);
void s3d_engine_destroy(s3d_engine *e);
const char *s3d_engine_error(const s3d_engine *e);
const char *s3d_last_create_error(void);
int s3d_engine_device(const s3d_engine *e);
/* Launch on a caller-owned cudaStream_t (e.g. torch's current stream) so the
 * caller's CUDA events bracket the kernels.  NULL restores the engine's own. */
int s3d_engine_set_stream(s3d_engine *e, void *cuda_stream);
void *s3d_engine_stream(const s3d_engine *e);
int s3d_engine_sync(s3d_engine *e);
/* Number of kernels this engine has launched since creation (bench: gpu_launches). */
long long s3d_engine_launch_count(const s3d_engine *e);

/* Icosahedron table exactly as init_geometry builds it (sift.c:215-326):
 * v[20][3][3] vertex vectors (after the reference's vertex swap), idx[20][3]. */
int s3d_set_mesh(s3d_engine *e, const float *v, const int *idx);

/* ---- pyramid (replaces resize_Pyramid imutil.c:3858, make_gss imutil.c:3752) --- */
/* gpyr: num_octaves*(K+3) levels, dog: num_octaves*(K+2) levels, octave-major,
 * first level s = -1.  Allocates HBM for every level. */
int s3d_pyramid_resize(s3d_engine *e, int num_octaves, int num_kp_levels,
                       const s3d_geom *gpyr, const s3d_geom *dog);
/* first: sigma_n -> s(0,-1); octave[j]: level j-1 -> j (sift.c:1012,1020). */
int s3d_pyramid_filters(s3d_engine *e, const s3d_filter *first, const s3d_filter *octave,
                        int num_octave_filters);

/* im_copy_data (imutil.c:1895): host image with element strides -> HBM. */
int s3d_image_upload(s3d_engine *e, const float *host, int nx, int ny, int nz, size_t xs,
                     size_t ys, size_t zs);
/* Same, source already resident in HBM (contiguous). */
int s3d_image_from_device(s3d_engine *e, const float *dev, int nx, int ny, int nz);

/* im_scale (imutil.c:1977) + build_gpyr (sift.c:989) + build_dog (sift.c:1052). */
int s3d_build_pyramid(s3d_engine *e);
/* detect_extrema (sift.c:1074): candidates stay on the device, in scan order. */
int s3d_detect_extrema(s3d_engine *e, double peak_thresh, int *num_candidates);
/* assign_orientations (sift.c:1264): rejects + stable compaction. */
int s3d_assign_orientations(s3d_engine *e, double corner_thresh, int *num_keypoints);
int s3d_candidates_download(s3d_engine *e, s3d_keypoint *out, int cap);
/* Counts left by the last s3d_detect_extrema / s3d_assign_orientations (num in detect_extrema,
 * sift.c:1197; kp->slab.num after assign_orientations, sift.c:1306-1324). */
int s3d_num_candidates(const s3d_engine *e);
int s3d_num_keypoints(const s3d_engine *e);
int s3d_keypoints_download(s3d_engine *e, s3d_keypoint *out, int cap);

/* _SIFT3D_extract_descriptors (sift.c:2207) on the resident Gaussian pyramid.
 * host_desc: n records of S3D_DESC_STRIDE bytes (the SIFT3D_Descriptor layout). */
int s3d_extract_descriptors(s3d_engine *e, const s3d_keypoint *kp, int n, void *host_desc);
/* Same with device-resident keypoints/descriptors (bench: no PCIe in the timed region). */
int s3d_extract_descriptors_device(s3d_engine *e, const s3d_keypoint *dev_kp, int n,
                                   void *dev_desc);
/* Keypoints left on the device by s3d_assign_orientations (for the above). */
const s3d_keypoint *s3d_device_keypoints(const s3d_engine *e);

/* One-level pyramid for the raw-image entry points (sift.c:2131-2195, 1534-1604):
 * uploads the image, applies smooth_scale_raw_input (sift.c:1978-2006) and leaves
 * the result as level (o=0, s=0) with absolute scale `scale`. */
int s3d_single_level(s3d_engine *e, const float *host, int nx, int ny, int nz, size_t xs, size_t ys,
                     size_t zs, const double units[3], double scale, const s3d_filter *smooth);
/* assign_eig_ori (sift.c:1354) on host-supplied keypoints: sigma = sig_fctr * kp.sd.
 * Fills kp[i].R; ok[i] = 0 where the reference returns REJECT (incl. conf < thresh);
 * conf[i] = corner score (0 on reject before the score exists). */
int s3d_orient_keypoints(s3d_engine *e, s3d_keypoint *kp, int n, double sig_fctr,
                         double corner_thresh, double *conf, unsigned char *ok);

/* SIFT3D_extract_dense_descriptors, dense_rotate = 0 (sift.c:2354-2496).
 * smooth: sigma_n -> sigma0 filter; window: the 12-channel blur filter;
 * units: of the input image; desc_units: of the caller's `desc` Image (the
 * reference blurs the channel image in THOSE units, sift.c:2451).
 * host_out: nx*ny*nz*12 floats, channel-interleaved. */
int s3d_dense_descriptors(s3d_engine *e, const float *host_in, int nx, int ny, int nz, size_t xs,
                          size_t ys, size_t zs, const double units[3],
                          const double desc_units[3], const s3d_filter *smooth,
                          const s3d_filter *window, float *host_out);

/* CUDA-event times of the last s3d_dense_descriptors call on the engine's stream:
 * ms[0] upload, ms[1] kernels (smoothing, channel image, 12-channel blur, post-processing),
 * ms[2] download.  -1 when no call has completed. */
int s3d_dense_last_timing(const s3d_engine *e, double ms[3]);

/* SIFT3D_extract_dense_descriptors, dense_rotate = 1 (sift.c:2521-2588, :2295-2343):
 * per voxel, assign_orientation_thresh with sigma = ori_sigma (identity when rejected),
 * then one 12-bin histogram over a sphere of radius 2 * desc_sigma, gradients rotated by
 * R^T; post-processing as in the no-rotate path.  host_out: nx*ny*nz*12 floats. */
int s3d_dense_descriptors_rotate(s3d_engine *e, const float *host_in, int nx, int ny, int nz,
                                 size_t xs, size_t ys, size_t zs, const double units[3],
                                 const s3d_filter *smooth, double ori_sigma, double desc_sigma,
                                 double corner_thresh, float *host_out);

/* Host copy of a pyramid level (which: 0 = Gaussian, 1 = DoG). */
int s3d_level_download(s3d_engine *e, int which, int o, int s, float *host_dst);
int s3d_pyramid_copy(s3d_engine *dst, const s3d_engine *src); /* copy_Pyramid, imutil.c:3995 */
int s3d_num_octaves(const s3d_engine *e);

/* ---- Z-slab tiling of ONE volume over several GPUs (BASELINE.json configs[4]) ------
 * No reference counterpart (the reference is one process in host memory); the contract is
 * its RESULT on the whole volume: identical pyramid bits, keypoints (order included) and
 * descriptors.  Rank r owns the octave-0 planes [zsplit[r], zsplit[r+1]) and, per octave,
 * the planes whose 2x-decimation source it owns (im_downsample_2x, imutil.c:1742-1768).
 * Communication: all-reduce(max) of max|image| (im_scale, imutil.c:1983) and max|DoG|
 * (sift.c:1161-1169), and halo planes of every Gaussian level (convolve_sep_gen z reach,
 * imutil.c:2274-2393; the orientation/descriptor windows, sift.c:1354, :1834). */
typedef struct s3d_comm s3d_comm;
/* in-process transport: one host thread per rank (any mix of devices, incl. all on one) */
void *s3d_local_world_create(int nranks);
void s3d_local_world_destroy(void *world);
int s3d_comm_create_local(s3d_comm **out, void *world, int rank, int device);
/* NCCL transport: one process per GPU; rank 0 creates the id and the caller broadcasts it
 * (e.g. torch.distributed); libnccl.so.2 is dlopen'ed ($SIFT3D_NCCL_LIB overrides) */
int s3d_nccl_unique_id(unsigned char id[128]);
int s3d_comm_create_nccl(s3d_comm **out, int rank, int nranks, const unsigned char id[128],
                         int device);
void s3d_comm_destroy(s3d_comm *c);
int s3d_comm_rank(const s3d_comm *c);
int s3d_comm_size(const s3d_comm *c);
const char *s3d_comm_error(const s3d_comm *c);
int s3d_comm_allreduce_max_u32(s3d_comm *c, s3d_engine *e, unsigned *dev, int n);
/* Host logic: planes owned per rank and octave, own[(r*num_octaves + o)*2 + {0,1}]. */
int s3d_slab_plan(int nranks, int num_octaves, int nz0, const int *zsplit, int *own);
/* Host logic: halo transfers of one level of octave o (NZ planes) for halo width h, as rows
 * {kind: 0 send / 1 recv, peer, z0, z1} in the order both ends of a pair use. */
int s3d_slab_halo_plan(int nranks, int num_octaves, const int *own, int o, int NZ, int h, int rank,
                       int *out, int cap);
/* Like s3d_pyramid_resize, with the GLOBAL level geometry; call s3d_pyramid_filters first.
 * Afterwards s3d_build_pyramid / s3d_detect_extrema / s3d_assign_orientations /
 * s3d_extract_descriptors work on the slab; keypoint coordinates are global. */
int s3d_slab_pyramid_resize(s3d_engine *e, s3d_comm *comm, int num_octaves, int num_kp_levels,
                            const s3d_geom *gpyr, const s3d_geom *dog, const int *zsplit);
/* The rank's owned planes of the input, nx*ny*(zsplit[r+1]-zsplit[r]) voxels. */
int s3d_slab_image_upload(s3d_engine *e, const float *host, size_t xs, size_t ys, size_t zs);
int s3d_slab_image_from_device(s3d_engine *e, const float *dev);
/* info = {own0, own1, lo, hi, NZ, halo} of octave o: owned planes, planes held, global count. */
int s3d_slab_info(const s3d_engine *e, int o, int info[6]);
/* Communication of the last s3d_build_pyramid on a slab engine: out = {halo bytes sent, halo
 * bytes received, exchanges, ms in halo exchanges, ms in all-reduces}; the two times are CUDA-event
 * intervals on the engine's stream (they include waiting for the slower neighbour) and need
 * s3d_set_option(e, "slab_timing", 1), else -1. */
int s3d_slab_stats(s3d_engine *e, double out[5]);

/* ---- descriptor matching -------------------------------------------
 * SIFT3D_nn_match / match_desc (sift.c:2840-2969): nearest + second-nearest by f64 SSD over
 * the 768 histogram values (accumulated in the reference's order, bit-identical), ratio test
 * ssd_best / ssd_nearest > nn_thresh^2, forward-backward consistency.  d1/d2: records of
 * S3D_DESC_STRIDE bytes; matches[i] = index into d2 or -1. */
int s3d_nn_match(s3d_engine *e, const void *host_d1, int n1, const void *host_d2, int n2,
                 float nn_thresh, int *host_matches);
int s3d_nn_match_device(s3d_engine *e, const void *dev_d1, int n1, const void *dev_d2, int n2,
                        float nn_thresh, int *dev_matches);

/* ---- resampling ------------------------------------------------------
 * im_inv_transform (imutil.c:2040-2081) for an Affine tform (apply_Affine_xyz,
 * imutil.c:2651): dst[x,y,z,c] = resample(src, A * [x y z 1]^T), A row-major 3x4 in f64;
 * interp 0 = LINEAR (resample_linear, imutil.c:2085: trilinear in f64, zero outside
 * [0, n-1]; bit-identical to the CPU), 1 = LANCZOS2 (resample_lanczos2, imutil.c:2127).
 * Volumes are contiguous, channel-interleaved (im_default_stride, imutil.c:1453). */
int s3d_resample_affine(s3d_engine *e, const float *host_src, int nx, int ny, int nz, int nc,
                        const double A[12], int interp, float *host_dst, int dnx, int dny, int dnz);
int s3d_resample_affine_device(s3d_engine *e, const float *dev_src, int nx, int ny, int nz, int nc,
                               const double A[12], int interp, float *dev_dst, int dnx, int dny,
                               int dnz);

/* ---- kernel-level entry (tests / bench roofline): device pointers ------------ */
/* apply_Sep_FIR_filter (imutil.c:3459): x, y, z passes, nc interleaved channels. */
int s3d_blur_device(s3d_engine *e, const float *dev_src, float *dev_dst, int nx, int ny, int nz,
                    int nc, const float *taps, int width, double unit, const double units[3]);
/* 0 = auto (fused fast path when eligible), 1 = force the generic per-axis path. */
int s3d_set_blur_mode(s3d_engine *e, int mode);
/* Debug / tuning switches (name: default -- meaning).  None is needed in production; the A/B
 * records under profiles/ were taken with them.
 *   icos_fast: 1, blur_mode: 0, desc_v1: 0 -- literal variants of the kernels
 *   desc_v2: 0 -- 1 = round 1's descriptor kernel k_descriptor2 instead of k_descriptor3
 *   desc_path: 0 -- 1..3 force one of the descriptor kernel's fixed-point fallback paths, +4 leaves
 *       its row intervals untrimmed (tests only)
 *   desc_occ: 4 -- CTAs per SM the descriptor kernel is compiled for (3)
 *   desc_pre: 0 -- 1 = the next voxel's gradient is fetched one trip ahead
 *   desc_norot: 0 -- 1 = no lane-dependent vertex order in k_descriptor2's scatter
 *   desc_streams: 2 -- s3d_extract_descriptors queues its chunks on two alternating compute
 *       streams, so that a chunk's draining CTAs overlap the next chunk; 1 = one stream
 *   desc_chunk: 4096 -- keypoints per chunk of that call (one launch + one D2H copy each)
 *   orient_v1: 0 -- 1 = the thread-per-candidate orientation kernel instead of the grouped one
 *   orient_batch: 4, orient_g: 8, orient_scalar: 0 -- variants of the orientation kernels
 *   dense_copy: 1 -- staged parallel copies to / from pageable host memory (0 = plain cudaMemcpy)
 *   copy_pipe: -1 -- how those staged copies run: 0 = one chunk at a time split over the host
 *       threads; 1 = every thread moves whole chunks through a ring of pipe_slots pinned slots of
 *       pipe_chunk_kb KB (0 = automatic: 16 x 2048) while the caller issues the DMAs and polls;
 *       2 = every thread incl. the caller owns a stream and two slots and runs its chunks alone
 *       (csrc/host_pipe.h); -1 = automatic: 1, or 2 when $LOCAL_WORLD_SIZE > 1, where a rank has
 *       too few host threads to spare one for polling
 *   blur_v1: 0 -- 1 = round 1's fused Gaussian k_blur_fused (LDG fill) also where the TMA-fed
 *       k_blur_tma is eligible
 *   blur_rpt4_hw: 3 -- widest filter half-width that takes k_blur_tma's 64 x 64 tile
 *   blur_w0 .. blur_w3 -- permille: per-plane cost of a left / right / top / bottom edge column of
 *       the fused blur relative to an interior one (balances the persistent CTAs' z ranges)
 *   blur_slabs: 0 -- z slabs the fused blur's work list is ordered by (0 = automatic)
 *   slab_timing: 0 -- 1 = CUDA-event times of the halo exchanges (s3d_slab_stats)
 *   blur_dbg, blur_flags -- per-CTA clocks of the last fused blur / experiment flags */
int s3d_set_option(s3d_engine *e, const char *name, int value);
/* Debug: per-CTA {start, end clock, SM id, steps} of the last fused blur (after option "blur_dbg"). */
int s3d_debug_read(s3d_engine *e, void *host, size_t bytes);
void *s3d_dev_alloc(s3d_engine *e, size_t bytes);
void s3d_dev_free(s3d_engine *e, void *p);
int s3d_memcpy_h2d(s3d_engine *e, void *dev, const void *host, size_t bytes);
int s3d_memcpy_d2h(s3d_engine *e, void *host, const void *dev, size_t bytes);
/* Tests: host_src -> device -> host_dst through the paths the library takes for a caller's
 * PAGEABLE buffers (im_copy_data of the input Image, imutil.c:1895; the dense result written into
 * the caller's Image, sift.c:2375-2380): staged copies for >= 32 / 64 MB, see option "copy_pipe". */
int s3d_host_roundtrip(s3d_engine *e, const void *host_src, void *host_dst, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif
