// common.cuh -- shared declarations of the B200 SIFT3D device engine.
//
// All arithmetic that feeds a parity-checked result is written with explicit
// round-to-nearest intrinsics (__fmul_rn/__fadd_rn/...) and the whole library is
// compiled with -fmad=false: the reference is built for baseline x86-64 (no FMA),
// so every multiply and add there rounds separately.
#pragma once

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>  // header-only NVTX 3: ranges cost nothing unless a tool is attached

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/sift3d_cuda.h"

#define S3D_NUM_SMS_FALLBACK 148

struct TapSet {  // passed by value as a kernel parameter (lands in the constant bank)
    float t[S3D_MAX_TAPS];
    int width;
};

struct LevelDev {
    s3d_geom g;
    float *d = nullptr;
    size_t n() const { return (size_t)g.nx * g.ny * g.nz; }
};

struct Candidate {  // 16 bytes, device-side candidate / keypoint coordinate record
    short o, s;
    int x, y, z;
};

// per-face constants of the Moller-Trumbore test, derived once from the mesh in
// the reference's f32 operation order (cart2bary, sift.c:335-394)
struct FaceConst {
    float e1[3], e2[3], t[3], q[3];
    float e2q;
    int idx[3];
    float vmid[3];  // unit centroid direction (fast-path preselection only)
};

struct MeshDev {
    FaceConst f[20];
    float vert[12][3];  // unit vertex directions (fast-path preselection only)
};

struct SegTab {  // cached work decomposition of the fused blur for one volume size / z range
    int nx, ny, nz, grid;
    int zb, ze;  // output planes [zb, ze) (0, nz for a whole volume)
    int hw;      // filter half-width (halo slivers and edge-column weights depend on it)
    int ty;      // tile height the table was cut for (32: k_blur_fused / k_blur_tma<.,2>, 64: k_blur_tma<.,4>)
    void *d;
    size_t nseg;
};

// Z-slab tiling (slab.cu): the planes of one octave this rank owns / holds.
struct SlabOct {
    int NZ = 0;          // global plane count of the octave
    int own0 = 0, own1 = 0;  // owned planes [own0, own1), global indices
    int lo = 0, hi = 0;      // planes held locally [lo, hi) = owned + halo, clipped to [0, NZ)
};

struct s3d_comm;  // slab.cu

struct s3d_engine {
    int device = 0;
    int num_sms = S3D_NUM_SMS_FALLBACK;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;  // D2H of descriptor chunks behind the kernel
    cudaStream_t aux_stream = nullptr;   // second compute stream of the chunked descriptor launches
    int opt_desc_streams = 2;            // host-buffer descriptor call: chunks alternate between 2 streams (1 = one)
    int opt_desc_chunk = 4096;           // keypoints per chunk of that call (one kernel launch + one D2H copy each)
    std::string err;
    long long launches = 0;
    int blur_mode = 0;
    int opt_icos_fast = 1;
    int opt_desc_v1 = 0;
    int opt_desc_pre = 0;  // 1: k_descriptor3 fetches the next voxel's gradient one trip ahead
    int opt_desc_v2 = 0;  // 1: k_descriptor2 (raster-order rows) instead of k_descriptor3 (cell-owner lanes)
    int opt_orient_scalar = 0;  // 1: k_orient_group reads the level (7 loads) instead of the gradient volume
    int opt_orient_g = 8;       // lanes per candidate of k_orient_group (8, 16 or 32)
    int opt_orient_v1 = 0;  // 1: thread-per-candidate k_orient instead of k_orient_group (A/B, tests)
    int ori_max_twx = 0;       // widest weight-table row of the current orientation tables
    int opt_orient_batch = 4;  // voxels k_orient fetches ahead (4, or 8 = line-aligned batches)
    int opt_desc_norot = 0;  // 1: no lane-dependent vertex order in k_descriptor2 (A/B only)
    int opt_desc_occ = 4;   // CTAs per SM k_descriptor2 is compiled for (3 or 4)
    int opt_desc_path = 0;  // test hook: force a fixed-point path of k_descriptor2 (0 = automatic)
    double blur_w[4] = {1.05, 1.10, 1.05, 1.10};  // per-plane cost of edge columns (left,right,top,bottom)
    int opt_blur_slabs = 0;    // z slabs of the fused blur's work list (0 = automatic; A/B only)
    bool blur_w_user = false;  // set through options blur_w0..3: then used for every filter width
    int opt_blur_flags = 0;  // timing experiments only (results wrong when non-zero)
    int opt_blur_v1 = 0;     // 1: k_blur_fused (round 1, LDG fill) instead of k_blur_tma (A/B, tests)
    int opt_blur_rpt4_hw = 3;  // widest half-width that takes the 64 x 64 tile of k_blur_tma (w = 9 measured slower with it)

    // pyramid
    int noct = 0, K = 0, nlev_g = 0, nlev_d = 0;
    int first_level = -1;  // index of the first level of an octave (-1 for detector pyramids)
    std::vector<LevelDev> g, dog;
    TapSet first_taps{};
    std::vector<TapSet> oct_taps;
    float *im = nullptr;  // scaled input copy (sift3d->im)
    int im_nx = 0, im_ny = 0, im_nz = 0;
    size_t im_cap = 0;
    float *scratch[2] = {nullptr, nullptr};
    size_t scratch_cap = 0;
    unsigned *d_scalars = nullptr;  // [0] = image max bits, [1 + o*nlev_d + (s+1)] = dogmax bits
    int n_scalars = 0;

    // extrema / keypoints
    Candidate *d_cand = nullptr;
    int cand_cap = 0;
    int ncand = 0;
    unsigned *d_mask = nullptr;  // bit masks, per keypoint level of the current octave
    size_t mask_cap = 0;
    int *d_blockcnt = nullptr;
    size_t blockcnt_cap = 0;
    int *d_counter = nullptr;  // [0] running candidate total, [1] keypoint total
    s3d_keypoint *d_kp_all = nullptr;  // candidates as keypoint records (R filled by k_orient)
    double *d_conf = nullptr;
    unsigned char *d_ok = nullptr;
    int *d_pos = nullptr;
    s3d_keypoint *d_kp = nullptr;  // compacted keypoints
    int kp_cap = 0;
    int nkp = 0;

    // gradient volumes of the keypoint levels (keypoint.cu): float4 {gx, gy, gz, 0} per voxel,
    // the central differences both the orientation and the descriptor kernels need -- one
    // 16-byte gather per voxel instead of six 4-byte ones.  Optional (nullptr: scalar path).
    std::vector<float4 *> grad;        // per gpyr level
    std::vector<size_t> grad_cap;      // voxels allocated
    std::vector<float4 *> grad_uploaded;  // what d_level_gptrs currently holds
    float4 **d_level_gptrs = nullptr;  // device table (nullptr entries allowed)
    bool grad_valid = false;           // volumes match the current pyramid contents

    // orientation window-weight tables (keypoint.cu)
    float *d_ori_pool = nullptr;
    size_t ori_pool_cap = 0;
    void *d_ori_tabs = nullptr;
    size_t ori_tabs_cap = 0;
    void *d_ori_lists = nullptr;   // in-sphere offset lists (same offsets as the pool)
    bool ori_lists_ok = false;
    std::vector<double> ori_key;   // scales / units the tables were built for
    int ori_key_rc = 1;

    // dense-descriptor buffers (raw, smoothed, 12-channel temp, 12-channel result), kept between
    // calls, and the pinned staging ring of the pageable-destination download
    float *dense_buf[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t dense_cap[4] = {0, 0, 0, 0};
    void *stage[2] = {nullptr, nullptr};
    size_t stage_cap = 0;
    cudaEvent_t dense_ev[4] = {nullptr, nullptr, nullptr, nullptr};
    double dense_ms[3] = {-1.0, -1.0, -1.0};  // upload, kernels, download of the last dense call
    int opt_dense_copy = 1;  // 1 = staged parallel download, 0 = plain cudaMemcpy into pageable memory
    // chunk pipeline between pageable host memory and the device (engine.cu: PipeJob)
    void *pipe_buf = nullptr;
    size_t pipe_cap = 0;
    std::vector<cudaEvent_t> pipe_ev;
    std::vector<cudaStream_t> pipe_streams;  // copy_pipe = 2: one per participant
    int opt_copy_pipe = -1;       // staged copies >= 32 MB: 0 = round 2's chunk-at-a-time copy, 1 = whole-chunk workers + a polling
                                  // caller, 2 = poller-free (a stream per host thread); -1 = automatic: 1, or 2 when ranks share the host
    int opt_pipe_chunk_kb = 0;    // 0 = automatic (2048; 1024 with several ranks on the host)
    int opt_pipe_slots = 0;       // 0 = automatic (16; 8 with several ranks on the host)

    // descriptor scratch
    s3d_keypoint *d_kp_in = nullptr;
    int kp_in_cap = 0;
    unsigned char *d_desc = nullptr;
    size_t desc_cap = 0;

    // Z-slab tiling: empty `slab` = the engine holds whole volumes
    s3d_comm *comm = nullptr;        // not owned
    std::vector<SlabOct> slab;       // per octave
    std::vector<int> zsplit;         // octave-0 plane split over the ranks (nranks + 1 entries)
    std::vector<int> slab_own;       // [rank][octave] -> own0, own1 of every rank
    std::vector<int> slab_need;      // [octave][gpyr level] -> halo planes its consumers read
    std::vector<s3d_geom> slab_g;    // global geometry of the Gaussian levels
    int slab_halo = 0;               // halo planes kept around the owned range of every level
    int *d_level_zoff = nullptr;     // per gpyr level: global z of local plane 0 (0 when not tiled)
    // statistics of the last s3d_slab_build_pyramid (s3d_slab_stats): halo bytes sent / received,
    // number of exchanges, and -- with option "slab_timing" -- CUDA-event time spent in them
    double slab_sent = 0, slab_recv = 0, slab_xchg_ms = 0, slab_allreduce_ms = 0;
    int slab_nxchg = 0;
    int opt_slab_timing = 0;
    std::vector<cudaEvent_t> slab_ev;  // event pool of the timed exchanges
    std::vector<int> slab_ev_kind;     // per event pair: 0 halo exchange, 1 all-reduce

    std::vector<SegTab> segtabs;
    long long *d_blur_dbg = nullptr;  // debug: per-CTA clocks of the last fused blur

    MeshDev *d_mesh = nullptr;
    bool have_mesh = false;
    float **d_level_ptrs = nullptr;  // device table of gpyr level pointers
    int *d_level_dims = nullptr;     // nx,ny,nz per gpyr level
    float *d_level_units = nullptr;  // (float)ux,uy,uz per gpyr level
    double *d_level_scales = nullptr;  // Image.s per gpyr level
};

// NVTX range per pipeline stage (nsys / ncu --nvtx): "s3d:upload", "s3d:pyramid", "s3d:extrema",
// "s3d:orientation", "s3d:descriptors", "s3d:dense", "s3d:halo_exchange"
struct S3dRange {
    explicit S3dRange(const char *name) { nvtxRangePushA(name); }
    ~S3dRange() { nvtxRangePop(); }
};

int s3d_fail(s3d_engine *e, const char *what, cudaError_t ce, const char *file, int line);
void s3d_free_pyramid(s3d_engine *e);

#define S3D_CUDA(e, call)                                                     \
    do {                                                                      \
        cudaError_t _ce = (call);                                             \
        if (_ce != cudaSuccess) return s3d_fail((e), #call, _ce, __FILE__, __LINE__); \
    } while (0)

#define S3D_LAUNCH_CHECK(e)                                                   \
    do {                                                                      \
        (e)->launches++;                                                      \
        cudaError_t _ce = cudaGetLastError();                                 \
        if (_ce != cudaSuccess) return s3d_fail((e), "kernel launch", _ce, __FILE__, __LINE__); \
    } while (0)

// ---- launchers implemented in pyramid.cu ------------------------------------
int s3d_k_max_abs(s3d_engine *e, const float *x, size_t n, unsigned *d_bits);
int s3d_k_scale(s3d_engine *e, const float *src, float *dst, size_t n, const unsigned *d_bits);
int s3d_k_blur(s3d_engine *e, const float *src, float *dst, int nx, int ny, int nz, int nc,
               const TapSet &taps, const float uf[3]);
// same, output restricted to planes [zb, ze) of a buffer that is the window
// [gz0, gz0 + nz) of a volume of nz_glob planes (src planes within the filter's z reach of
// the range must be valid; the mirror rules apply at the true volume ends only)
int s3d_k_blur_zrange(s3d_engine *e, const float *src, float *dst, int nx, int ny, int nz,
                      const TapSet &taps, const float uf[3], int zb, int ze, int gz0, int nz_glob);
// planes beyond an output plane that a z pass with these taps reads (incl. the lerp partner)
int s3d_blur_z_reach(const TapSet &taps, float ufz);
int s3d_k_decimate(s3d_engine *e, const float *src, int sx, int sy, int sz, float *dst, int dx,
                   int dy, int dz);
int s3d_k_dog(s3d_engine *e, const float *a, const float *b, float *d, size_t n,
              unsigned *d_maxbits);
int s3d_k_dog_octave(s3d_engine *e, int o);  // all DoG levels of an octave at once; 1 = not applicable
int s3d_k_extrema_octave(s3d_engine *e, int o, float peak_thresh_dummy, double peak_thresh);
// scan only local planes [zl0, zl0 + nzs) of the octave's DoG buffers as if they were a whole
// volume (its first and last plane are neighbours only); emitted z = local-sub z + zbase
int s3d_k_extrema_range(s3d_engine *e, int o, double peak_thresh, int zl0, int nzs, int zbase);
int s3d_ensure_scratch(s3d_engine *e, size_t elems);

// ---- launchers implemented in keypoint.cu -----------------------------------
int s3d_k_orientations(s3d_engine *e, double corner_thresh);
int s3d_k_orient_list(s3d_engine *e, s3d_keypoint *d_kp, int n, double sig_fctr,
                      double corner_thresh, unsigned char *d_ok, double *d_conf);
// kp_levels_only: every keypoint sits on a keypoint level s = 0..K-1 of the resident pyramid and
// its window passes s3d_desc_window_fine (then the cell-owner kernel k_descriptor3 is used)
int s3d_k_descriptors(s3d_engine *e, const s3d_keypoint *d_kp, int n, unsigned char *d_out,
                      int kp_levels_only);
int s3d_k_dense(s3d_engine *e, const float *d_smooth, const float *d_raw, int nx, int ny, int nz,
                const float inv_units[3], float *d_temp12);
int s3d_k_dense_rotate(s3d_engine *e, const float *d_smooth, int nx, int ny, int nz,
                       const float units[3], double ori_sigma, double desc_sigma,
                       double corner_thresh, float *d_out12);
int s3d_k_dense_post(s3d_engine *e, float *d_desc12, const float *d_raw, size_t nvox);
// Whether the one-word fixed-point histogram of k_descriptor3 keeps >= 18 bits below the largest
// gradient of the window for a keypoint of scale sd on a level with these units: the sum of the
// trilinear weights of a bin (hist_width + voxel diagonal)^3 / voxel volume, hist_width = 5 sd
// (sift.c:1845-1850), must stay below 8192 voxels.  Larger windows (user keypoints at coarse
// scales on a fine level) take k_descriptor2 with its two-word bins.
inline bool s3d_desc_window_fine(double sd, double ux, double uy, double uz)
{
    const double hw = sd * 7.071067812 * 2.0 / 1.4142135623730951 / 2.0;
    const double side = hw + sqrt(ux * ux + uy * uy + uz * uz);
    return side * side * side / (ux * uy * uz) <= 8192.0;
}
int s3d_upload_mesh(s3d_engine *e, const float *v, const int *idx);
int s3d_gradients_prepare(s3d_engine *e);  // best effort: 0 also when memory is short
void s3d_gradients_free(s3d_engine *e);
