// engine.cu -- the C-ABI of libsift3d_cuda.so (see include/sift3d_cuda.h).
//
// One engine = one CUDA device + one stream + the HBM-resident pyramids of one
// SIFT3D object.  Host code (plain C, sift3d_b200/host/) owns parameters, filter
// design and the reference-compatible structs; everything per-voxel happens here.
#include "common.cuh"
#include "host_pipe.h"

#include <sched.h>

#include <atomic>
#include <ctime>
#include <functional>
#include <memory>
#include <thread>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <mutex>

static thread_local std::string g_create_err;

int s3d_fail(s3d_engine *e, const char *what, cudaError_t ce, const char *file, int line)
{
    char buf[512];
    snprintf(buf, sizeof(buf), "sift3d_cuda: %s failed: %s (%s:%d)", what,
             ce == cudaSuccess ? "invalid state/argument" : cudaGetErrorString(ce), file, line);
    if (e) e->err = buf;
    g_create_err = buf;
    fprintf(stderr, "%s\n", buf);
    return -1;
}

int s3d_pack_candidates(s3d_engine *e, s3d_keypoint *d_out);  // keypoint.cu
int s3d_slab_build_pyramid(s3d_engine *e);                     // slab.cu
int s3d_slab_extrema_octave(s3d_engine *e, int o, double peak_thresh);

void s3d_free_pyramid(s3d_engine *e)
{
    for (auto &l : e->g)
        if (l.d) cudaFree(l.d);
    for (auto &l : e->dog)
        if (l.d) cudaFree(l.d);
    e->g.clear();
    e->dog.clear();
    if (e->d_level_ptrs) cudaFree(e->d_level_ptrs);
    if (e->d_level_dims) cudaFree(e->d_level_dims);
    if (e->d_level_units) cudaFree(e->d_level_units);
    if (e->d_level_scales) cudaFree(e->d_level_scales);
    if (e->d_level_zoff) cudaFree(e->d_level_zoff);
    if (e->d_scalars) cudaFree(e->d_scalars);
    e->d_level_ptrs = nullptr;
    e->d_level_dims = nullptr;
    e->d_level_units = nullptr;
    e->d_level_scales = nullptr;
    e->d_level_zoff = nullptr;
    e->d_scalars = nullptr;
    e->noct = 0;
    s3d_gradients_free(e);
    e->slab.clear();
    e->slab_own.clear();
    e->slab_need.clear();
    e->slab_g.clear();
    e->zsplit.clear();
    e->comm = nullptr;
}

namespace {

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev)
    {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
    }
    ~DeviceGuard()
    {
        int cur = -1;
        cudaGetDevice(&cur);
        if (prev >= 0 && cur != prev) cudaSetDevice(prev);
    }
};

void free_pyramid(s3d_engine *e) { s3d_free_pyramid(e); }

int ensure_cand(s3d_engine *e, int cap)
{
    if (cap <= e->cand_cap) return 0;
    if (e->d_cand) cudaFree(e->d_cand);
    if (e->d_kp_all) cudaFree(e->d_kp_all);
    if (e->d_ok) cudaFree(e->d_ok);
    if (e->d_pos) cudaFree(e->d_pos);
    if (e->d_kp) cudaFree(e->d_kp);
    e->d_cand = nullptr;
    e->d_kp_all = nullptr;
    e->d_ok = nullptr;
    e->d_pos = nullptr;
    e->d_kp = nullptr;
    e->cand_cap = 0;
    S3D_CUDA(e, cudaMalloc(&e->d_cand, (size_t)cap * sizeof(Candidate)));
    S3D_CUDA(e, cudaMalloc(&e->d_kp_all, (size_t)cap * sizeof(s3d_keypoint)));
    S3D_CUDA(e, cudaMalloc(&e->d_ok, (size_t)cap));
    S3D_CUDA(e, cudaMalloc(&e->d_pos, (size_t)cap * sizeof(int)));
    S3D_CUDA(e, cudaMalloc(&e->d_kp, (size_t)cap * sizeof(s3d_keypoint)));
    e->cand_cap = cap;
    e->kp_cap = cap;
    return 0;
}

int to_tapset(s3d_engine *e, const s3d_filter *f, TapSet &t)
{
    if (!f || !f->taps || f->width < 1 || f->width > S3D_MAX_TAPS || !(f->width & 1))
        return s3d_fail(e, "filter width (odd, <= S3D_MAX_TAPS)", cudaSuccess, __FILE__, __LINE__);
    memset(&t, 0, sizeof(t));
    memcpy(t.t, f->taps, f->width * sizeof(float));
    t.width = f->width;
    return 0;
}

void level_uf(const s3d_geom &g, double unit, float uf[3])
{  // unit_factor = unit / units[dim], narrowed to f32 (imutil.c:2288-2289)
    uf[0] = (float)(unit / g.ux);
    uf[1] = (float)(unit / g.uy);
    uf[2] = (float)(unit / g.uz);
}

}  // namespace

extern "C" {

const char *s3d_last_create_error(void) { return g_create_err.c_str(); }

int s3d_engine_create(s3d_engine **out, int device)
{
    *out = nullptr;
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) {
        // No CPU fallback by design: the accelerated path IS the product.
        s3d_fail(nullptr, "cudaGetDeviceCount (no usable CUDA device; there is no CPU fallback)",
                 ce == cudaSuccess ? cudaErrorNoDevice : ce, __FILE__, __LINE__);
        return -1;
    }
    if (device < 0) {
        const char *env = getenv("SIFT3D_CUDA_DEVICE");
        if (!env) env = getenv("LOCAL_RANK");
        device = env ? atoi(env) : 0;
        if (device < 0 || device >= ndev) device = device % ndev;
    }
    if (device >= ndev) {
        s3d_fail(nullptr, "device index out of range", cudaErrorInvalidDevice, __FILE__, __LINE__);
        return -1;
    }
    s3d_engine *e = new s3d_engine();
    e->device = device;
    DeviceGuard guard(device);
    cudaDeviceProp prop;
    if ((ce = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
        s3d_fail(nullptr, "cudaGetDeviceProperties", ce, __FILE__, __LINE__);
        delete e;
        return -1;
    }
    e->num_sms = prop.multiProcessorCount;
    if ((ce = cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking)) != cudaSuccess) {
        s3d_fail(nullptr, "cudaStreamCreate", ce, __FILE__, __LINE__);
        delete e;
        return -1;
    }
    e->stream = e->own_stream;
    // tuning / A-B switches for the tools (same as s3d_set_option)
    if (const char *v = getenv("S3D_BLUR_MODE")) e->blur_mode = atoi(v);
    if (const char *v = getenv("S3D_DENSE_COPY")) e->opt_dense_copy = atoi(v);
    if (const char *v = getenv("S3D_COPY_PIPE")) e->opt_copy_pipe = atoi(v);
    if (const char *v = getenv("S3D_DESC_STREAMS")) e->opt_desc_streams = atoi(v);
    if (const char *v = getenv("S3D_PIPE_CHUNK_KB")) e->opt_pipe_chunk_kb = atoi(v);
    if (const char *v = getenv("S3D_PIPE_SLOTS")) e->opt_pipe_slots = atoi(v);
    if (const char *v = getenv("S3D_DESC_OCC")) e->opt_desc_occ = atoi(v);
    if (const char *v = getenv("S3D_DESC_NOROT")) e->opt_desc_norot = atoi(v);
    if (const char *v = getenv("S3D_ORIENT_BATCH")) e->opt_orient_batch = atoi(v);
    if ((ce = cudaMalloc(&e->d_counter, 4 * sizeof(int))) != cudaSuccess) {
        s3d_fail(nullptr, "cudaMalloc", ce, __FILE__, __LINE__);
        cudaStreamDestroy(e->own_stream);
        delete e;
        return -1;
    }
    *out = e;
    return 0;
}

void s3d_engine_destroy(s3d_engine *e)
{
    if (!e) return;
    DeviceGuard guard(e->device);
    cudaStreamSynchronize(e->stream);
    free_pyramid(e);
    void *ptrs[] = {e->im,    e->scratch[0], e->scratch[1], e->d_cand,  e->d_mask, e->d_blockcnt,
                    e->d_counter, e->d_kp_all, e->d_ok,       e->d_pos,   e->d_kp,   e->d_kp_in,
                    e->d_desc, e->d_mesh, e->d_blur_dbg, e->d_ori_pool, e->d_ori_tabs, e->d_ori_lists};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    for (auto &t : e->segtabs)
        if (t.d) cudaFree(t.d);
    for (float *p : e->dense_buf)
        if (p) cudaFree(p);
    for (void *p : e->stage)
        if (p) cudaFreeHost(p);
    if (e->pipe_buf) cudaFreeHost(e->pipe_buf);
    for (cudaEvent_t ev : e->pipe_ev) cudaEventDestroy(ev);
    for (cudaStream_t st : e->pipe_streams) cudaStreamDestroy(st);
    for (cudaEvent_t ev : e->dense_ev)
        if (ev) cudaEventDestroy(ev);
    for (cudaEvent_t ev : e->slab_ev) cudaEventDestroy(ev);
    if (e->own_stream) cudaStreamDestroy(e->own_stream);
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    if (e->aux_stream) cudaStreamDestroy(e->aux_stream);
    delete e;
}

const char *s3d_engine_error(const s3d_engine *e) { return e ? e->err.c_str() : ""; }
int s3d_engine_device(const s3d_engine *e) { return e->device; }
long long s3d_engine_launch_count(const s3d_engine *e) { return e->launches; }
void *s3d_engine_stream(const s3d_engine *e) { return (void *)e->stream; }

int s3d_engine_set_stream(s3d_engine *e, void *cuda_stream)
{
    DeviceGuard guard(e->device);
    S3D_CUDA(e, cudaStreamSynchronize(e->stream));
    e->stream = cuda_stream ? (cudaStream_t)cuda_stream : e->own_stream;
    return 0;
}

int s3d_engine_sync(s3d_engine *e)
{
    DeviceGuard guard(e->device);
    S3D_CUDA(e, cudaStreamSynchronize(e->stream));
    return 0;
}

int s3d_set_blur_mode(s3d_engine *e, int mode)
{
    e->blur_mode = mode;
    return 0;
}

int s3d_set_option(s3d_engine *e, const char *name, int value)
{
    if (!strcmp(name, "icos_fast")) e->opt_icos_fast = value;
    else if (!strcmp(name, "blur_mode")) e->blur_mode = value;
    else if (!strcmp(name, "desc_v1")) e->opt_desc_v1 = value;
    else if (!strcmp(name, "desc_v2")) e->opt_desc_v2 = value;
    else if (!strcmp(name, "desc_pre")) e->opt_desc_pre = value;
    else if (!strcmp(name, "slab_timing")) e->opt_slab_timing = value;
    else if (!strcmp(name, "desc_path")) e->opt_desc_path = value & 7;
    else if (!strcmp(name, "blur_flags")) e->opt_blur_flags = value;
    else if (!strcmp(name, "blur_v1")) e->opt_blur_v1 = value;
    else if (!strcmp(name, "blur_rpt4_hw")) e->opt_blur_rpt4_hw = value;
    else if (!strcmp(name, "blur_slabs")) {
        e->opt_blur_slabs = value;
        cudaStreamSynchronize(e->stream);
        for (auto &t : e->segtabs)
            if (t.d) cudaFree(t.d);
        e->segtabs.clear();
    } else if (!strncmp(name, "blur_w", 6) && name[6] >= '0' && name[6] <= '3' && !name[7]) {
        // per-plane cost of an edge column relative to an interior one, in permille (left, right,
        // top, bottom): the persistent CTAs' z ranges are balanced with it; cached tables are dropped
        e->blur_w[name[6] - '0'] = value * 1e-3;
        e->blur_w_user = true;
        cudaStreamSynchronize(e->stream);
        for (auto &t : e->segtabs)
            if (t.d) cudaFree(t.d);
        e->segtabs.clear();
    }
    else if (!strcmp(name, "dense_copy")) e->opt_dense_copy = value;
    else if (!strcmp(name, "copy_pipe")) e->opt_copy_pipe = value;
    else if (!strcmp(name, "pipe_chunk_kb")) e->opt_pipe_chunk_kb = value;
    else if (!strcmp(name, "pipe_slots")) e->opt_pipe_slots = value;
    else if (!strcmp(name, "desc_occ")) e->opt_desc_occ = value;
    else if (!strcmp(name, "desc_streams")) e->opt_desc_streams = value;
    else if (!strcmp(name, "desc_chunk")) e->opt_desc_chunk = value;
    else if (!strcmp(name, "desc_norot")) e->opt_desc_norot = value;
    else if (!strcmp(name, "orient_batch")) e->opt_orient_batch = value;
    else if (!strcmp(name, "orient_v1")) e->opt_orient_v1 = value;
    else if (!strcmp(name, "orient_scalar")) e->opt_orient_scalar = value;
    else if (!strcmp(name, "orient_g")) e->opt_orient_g = value;
    else if (!strcmp(name, "blur_dbg")) {
        DeviceGuard guard(e->device);
        if (value && !e->d_blur_dbg) {
            S3D_CUDA(e, cudaMalloc(&e->d_blur_dbg, 4 * 1024 * sizeof(long long)));
            S3D_CUDA(e, cudaMemset(e->d_blur_dbg, 0, 4 * 1024 * sizeof(long long)));
        }
    }
    else return s3d_fail(e, "s3d_set_option: unknown option", cudaSuccess, __FILE__, __LINE__);
    return 0;
}

int s3d_set_mesh(s3d_engine *e, const float *v, const int *idx)
{
    DeviceGuard guard(e->device);
    return s3d_upload_mesh(e, v, idx);
}

int s3d_num_octaves(const s3d_engine *e) { return e->noct; }

int s3d_pyramid_resize(s3d_engine *e, int num_octaves, int num_kp_levels, const s3d_geom *gpyr,
                       const s3d_geom *dog)
{
    DeviceGuard guard(e->device);
    S3D_CUDA(e, cudaStreamSynchronize(e->stream));
    const int nlev_d = num_kp_levels + 2, nlev_g = num_kp_levels + 3;
    // keep the allocation when nothing changed
    bool same = e->slab.empty() && e->noct == num_octaves && e->K == num_kp_levels &&
                (int)e->g.size() == num_octaves * nlev_g;
    for (int i = 0; same && i < num_octaves * nlev_g; i++)
        same = e->g[i].g.nx == gpyr[i].nx && e->g[i].g.ny == gpyr[i].ny && e->g[i].g.nz == gpyr[i].nz;
    if (same) {
        for (int i = 0; i < num_octaves * nlev_g; i++) e->g[i].g = gpyr[i];
        for (int i = 0; i < num_octaves * nlev_d; i++) e->dog[i].g = dog[i];
    } else {
        free_pyramid(e);
        e->noct = num_octaves;
        e->K = num_kp_levels;
        e->nlev_g = nlev_g;
        e->nlev_d = nlev_d;
        e->first_level = -1;
        e->g.resize((size_t)num_octaves * nlev_g);
        e->dog.resize((size_t)num_octaves * nlev_d);
        // a failed allocation must not leave a half-built pyramid that the `same` fast path of
        // the next call would accept: free everything (noct back to 0) before reporting
        cudaError_t ce = cudaSuccess;
        for (int i = 0; ce == cudaSuccess && i < num_octaves * nlev_g; i++) {
            e->g[i].g = gpyr[i];
            ce = cudaMalloc(&e->g[i].d, e->g[i].n() * sizeof(float));
        }
        for (int i = 0; ce == cudaSuccess && i < num_octaves * nlev_d; i++) {
            e->dog[i].g = dog[i];
            ce = cudaMalloc(&e->dog[i].d, e->dog[i].n() * sizeof(float));
        }
        const int L = num_octaves * nlev_g;
        e->n_scalars = 1 + num_octaves * nlev_d;
        if (ce == cudaSuccess) ce = cudaMalloc(&e->d_level_ptrs, L * sizeof(float *));
        if (ce == cudaSuccess) ce = cudaMalloc(&e->d_level_dims, 3 * L * sizeof(int));
        if (ce == cudaSuccess) ce = cudaMalloc(&e->d_level_units, 3 * L * sizeof(float));
        if (ce == cudaSuccess) ce = cudaMalloc(&e->d_level_scales, L * sizeof(double));
        if (ce == cudaSuccess) ce = cudaMalloc(&e->d_scalars, e->n_scalars * sizeof(unsigned));
        if (ce != cudaSuccess) {
            free_pyramid(e);
            return s3d_fail(e, "s3d_pyramid_resize: cudaMalloc", ce, __FILE__, __LINE__);
        }
    }
    if (num_octaves == 0) return 0;
    const int L = num_octaves * nlev_g;
    std::vector<float *> ptrs(L);
    std::vector<int> dims(3 * L);
    std::vector<float> units(3 * L);
    std::vector<double> scales(L);
    for (int i = 0; i < L; i++) {
        ptrs[i] = e->g[i].d;
        dims[3 * i] = gpyr[i].nx;
        dims[3 * i + 1] = gpyr[i].ny;
        dims[3 * i + 2] = gpyr[i].nz;
        units[3 * i] = (float)gpyr[i].ux;
        units[3 * i + 1] = (float)gpyr[i].uy;
        units[3 * i + 2] = (float)gpyr[i].uz;
        scales[i] = gpyr[i].scale;
    }
    S3D_CUDA(e, cudaMemcpy(e->d_level_ptrs, ptrs.data(), L * sizeof(float *), cudaMemcpyHostToDevice));
    S3D_CUDA(e, cudaMemcpy(e->d_level_dims, dims.data(), 3 * L * sizeof(int), cudaMemcpyHostToDevice));
    S3D_CUDA(e, cudaMemcpy(e->d_level_units, units.data(), 3 * L * sizeof(float), cudaMemcpyHostToDevice));
    S3D_CUDA(e, cudaMemcpy(e->d_level_scales, scales.data(), L * sizeof(double), cudaMemcpyHostToDevice));
    return 0;
}

int s3d_pyramid_filters(s3d_engine *e, const s3d_filter *first, const s3d_filter *octave,
                        int num_octave_filters)
{
    if (to_tapset(e, first, e->first_taps)) return -1;
    e->oct_taps.resize(num_octave_filters);
    for (int i = 0; i < num_octave_filters; i++)
        if (to_tapset(e, &octave[i], e->oct_taps[i])) return -1;
    return 0;
}

static int ensure_im(s3d_engine *e, int nx, int ny, int nz)
{
    const size_t n = (size_t)nx * ny * nz;
    if (n > e->im_cap) {
        if (e->im) cudaFree(e->im);
        e->im = nullptr;
        e->im_cap = 0;
        S3D_CUDA(e, cudaMalloc(&e->im, n * sizeof(float)));
        e->im_cap = n;
    }
    e->im_nx = nx;
    e->im_ny = ny;
    e->im_nz = nz;
    return 0;
}

static int h2d_pageable(s3d_engine *e, void *dev, const void *host, size_t bytes);

static int upload_strided(s3d_engine *e, float *dev, const float *host, int nx, int ny, int nz,
                          size_t xs, size_t ys, size_t zs)
{
    const size_t n = (size_t)nx * ny * nz;
    if (xs == 1 && ys == (size_t)nx && zs == (size_t)nx * ny) {
        // the caller's Image is usually malloc memory (pageable): staged parallel upload;
        // pinned / registered memory goes straight to the DMA engine
        cudaPointerAttributes at;
        const bool pageable = cudaPointerGetAttributes(&at, host) != cudaSuccess ||
                              at.type == cudaMemoryTypeUnregistered;
        cudaGetLastError();
        if (pageable) return h2d_pageable(e, dev, host, n * sizeof(float));
        S3D_CUDA(e, cudaMemcpyAsync(dev, host, n * sizeof(float), cudaMemcpyHostToDevice, e->stream));
        return 0;
    }
    if (xs == 1) {  // padded rows/planes: a 3-D strided copy
        cudaMemcpy3DParms p = {};
        p.srcPtr = make_cudaPitchedPtr((void *)host, ys * sizeof(float), nx, zs / ys);
        p.dstPtr = make_cudaPitchedPtr(dev, (size_t)nx * sizeof(float), nx, ny);
        p.extent = make_cudaExtent((size_t)nx * sizeof(float), ny, nz);
        p.kind = cudaMemcpyHostToDevice;
        if (zs % ys == 0) {
            S3D_CUDA(e, cudaMemcpy3DAsync(&p, e->stream));
            return 0;
        }
    }
    // arbitrary element strides (im_copy_data honours them, imutil.c:1913-1916): gather first
    std::vector<float> tmp(n);
    for (int z = 0; z < nz; z++)
        for (int y = 0; y < ny; y++)
            for (int x = 0; x < nx; x++)
                tmp[x + (size_t)nx * (y + (size_t)ny * z)] = host[x * xs + y * ys + z * zs];
    S3D_CUDA(e, cudaMemcpyAsync(dev, tmp.data(), n * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    S3D_CUDA(e, cudaStreamSynchronize(e->stream));
    return 0;
}

int s3d_image_upload(s3d_engine *e, const float *host, int nx, int ny, int nz, size_t xs,
                     size_t ys, size_t zs)
{
    S3dRange nvtx_range("s3d:upload");
    DeviceGuard guard(e->device);
    if (ensure_im(e, nx, ny, nz)) return -1;
    return upload_strided(e, e->im, host, nx, ny, nz, xs, ys, zs);
}

int s3d_image_from_device(s3d_engine *e, const float *dev, int nx, int ny, int nz)
{
    DeviceGuard guard(e->device);
    if (ensure_im(e, nx, ny, nz)) return -1;
    S3D_CUDA(e, cudaMemcpyAsync(e->im, dev, (size_t)nx * ny * nz * sizeof(float),
                                cudaMemcpyDeviceToDevice, e->stream));
    return 0;
}

int s3d_build_pyramid(s3d_engine *e)
{
    S3dRange nvtx_range("s3d:pyramid");
    DeviceGuard guard(e->device);
    if (e->noct < 1 || !e->im || (int)e->oct_taps.size() != e->nlev_g - 1)
        return s3d_fail(e, "s3d_build_pyramid: pyramid/filters/image not configured", cudaSuccess,
                        __FILE__, __LINE__);
    e->grad_valid = false;
    if (!e->slab.empty()) return s3d_slab_build_pyramid(e);
    const LevelDev &base = e->g[0];
    if (base.g.nx != e->im_nx || base.g.ny != e->im_ny || base.g.nz != e->im_nz)
        return s3d_fail(e, "s3d_build_pyramid: image/pyramid size mismatch", cudaSuccess, __FILE__,
                        __LINE__);
    const size_t n0 = base.n();
    S3D_CUDA(e, cudaMemsetAsync(e->d_scalars, 0, e->n_scalars * sizeof(unsigned), e->stream));
    // im_scale (imutil.c:1977-1991)
    if (s3d_k_max_abs(e, e->im, n0, e->d_scalars)) return -1;
    if (s3d_k_scale(e, e->im, e->im, n0, e->d_scalars)) return -1;
    // build_gpyr (sift.c:989-1050); every hot call uses unit = 1.0 (sift.c:1002)
    float uf[3];
    level_uf(base.g, 1.0, uf);
    if (s3d_k_blur(e, e->im, e->g[0].d, base.g.nx, base.g.ny, base.g.nz, 1, e->first_taps, uf))
        return -1;
    for (int o = 0; o < e->noct; o++) {
        for (int s = 0; s <= e->nlev_g - 2; s++) {
            const LevelDev &src = e->g[(size_t)o * e->nlev_g + s];  // level s-1
            LevelDev &dst = e->g[(size_t)o * e->nlev_g + s + 1];    // level s
            level_uf(src.g, 1.0, uf);
            if (s3d_k_blur(e, src.d, dst.d, src.g.nx, src.g.ny, src.g.nz, 1, e->oct_taps[s], uf))
                return -1;
        }
        if (o != e->noct - 1) {  // im_downsample_2x of level max(s_end-2, first) (sift.c:1029-1041)
            const int ds = std::max(e->nlev_g - 2 - 2, -1);
            const LevelDev &src = e->g[(size_t)o * e->nlev_g + ds + 1];
            LevelDev &dst = e->g[(size_t)(o + 1) * e->nlev_g];
            if (s3d_k_decimate(e, src.d, src.g.nx, src.g.ny, src.g.nz, dst.d, dst.g.nx, dst.g.ny,
                               dst.g.nz))
                return -1;
        }
    }
    // build_dog (sift.c:1052-1071) fused with the per-level max|DoG| (sift.c:1161-1166)
    for (int o = 0; o < e->noct; o++) {
        const int drc = e->blur_mode == 0 ? s3d_k_dog_octave(e, o) : 1;
        if (drc < 0) return -1;
        if (drc == 0) continue;
        for (int s = -1; s <= e->nlev_d - 2; s++) {
            const LevelDev &a = e->g[(size_t)o * e->nlev_g + s + 1];
            const LevelDev &b = e->g[(size_t)o * e->nlev_g + s + 2];
            LevelDev &d = e->dog[(size_t)o * e->nlev_d + s + 1];
            if (s3d_k_dog(e, a.d, b.d, d.d, a.n(), e->d_scalars + 1 + (size_t)o * e->nlev_d + s + 1))
                return -1;
        }
    }
    return 0;
}

int s3d_detect_extrema(s3d_engine *e, double peak_thresh, int *num_candidates)
{
    S3dRange nvtx_range("s3d:extrema");
    DeviceGuard guard(e->device);
    if (e->noct < 1)
        return s3d_fail(e, "s3d_detect_extrema: no pyramid", cudaSuccess, __FILE__, __LINE__);
    const size_t n0 = e->dog[0].n();
    int want_cap = (int)std::min<size_t>(std::max<size_t>(n0 / 64, (size_t)1 << 16), (size_t)1 << 28);
    if (e->cand_cap > want_cap) want_cap = e->cand_cap;
    for (int attempt = 0; attempt < 8; attempt++) {
        if (ensure_cand(e, want_cap)) return -1;
        S3D_CUDA(e, cudaMemsetAsync(e->d_counter, 0, 4 * sizeof(int), e->stream));
        for (int o = 0; o < e->noct; o++)
            if (e->slab.empty() ? s3d_k_extrema_octave(e, o, 0.0f, peak_thresh)
                                : s3d_slab_extrema_octave(e, o, peak_thresh))
                return -1;
        int total = 0;
        S3D_CUDA(e, cudaMemcpyAsync(&total, e->d_counter, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
        S3D_CUDA(e, cudaStreamSynchronize(e->stream));
        if (total <= e->cand_cap) {
            e->ncand = total;
            e->nkp = 0;
            if (num_candidates) *num_candidates = total;
            return 0;
        }
        want_cap = total + total / 8 + 1024;  // overflowed: grow and redo the (cheap) scan
    }
    return s3d_fail(e, "s3d_detect_extrema: candidate buffer", cudaSuccess, __FILE__, __LINE__);
}

int s3d_assign_orientations(s3d_engine *e, double corner_thresh, int *num_keypoints)
{
    S3dRange nvtx_range("s3d:orientation");
    DeviceGuard guard(e->device);
    if (s3d_k_orientations(e, corner_thresh)) return -1;
    int nkp = 0;
    if (e->ncand > 0) {
        S3D_CUDA(e, cudaMemcpyAsync(&nkp, e->d_counter + 1, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
        S3D_CUDA(e, cudaStreamSynchronize(e->stream));
    }
    e->nkp = nkp;
    if (num_keypoints) *num_keypoints = nkp;
    return 0;
}

int s3d_keypoints_download(s3d_engine *e, s3d_keypoint *out, int cap)
{
    DeviceGuard guard(e->device);
    const int n = std::min(cap, e->nkp);
    if (n <= 0) return 0;
    S3D_CUDA(e, cudaMemcpyAsync(out, e->d_kp, (size_t)n * sizeof(s3d_keypoint), cudaMemcpyDeviceToHost, e->stream));
    S3D_CUDA(e, cudaStreamSynchronize(e->stream));
    return 0;
}

int s3d_candidates_download(s3d_engine *e, s3d_keypoint *out, int cap)
{
    DeviceGuard guard(e->device);
    const int n = std::min(cap, e->ncand);
    if (n <= 0) return 0;
    // d_kp_all holds the candidates (with R of the last orientation pass, if any)
    if (e->nkp == 0 && s3d_pack_candidates(e, e->d_kp_all)) return -1;
    S3D_CUDA(e, cudaMemcpyAsync(out, e->d_kp_all, (size_t)n * sizeof(s3d_keypoint), cudaMemcpyDeviceToHost, e->stream));
    S3D_CUDA(e, cudaStreamSynchronize(e->stream));
    return 0;
}

const s3d_keypoint *s3d_device_keypoints(const s3d_engine *e) { return e->d_kp; }
int s3d_num_candidates(const s3d_engine *e) { return e->ncand; }
int s3d_num_keypoints(const s3d_engine *e) { return e->nkp; }

int s3d_extract_descriptors_device(s3d_engine *e, const s3d_keypoint *dev_kp, int n,
                                   void *dev_desc)
{
    S3dRange nvtx_range("s3d:descriptors");
    DeviceGuard guard(e->device);
    if (e->noct < 1)
        return s3d_fail(e, "s3d_extract_descriptors: no pyramid", cudaSuccess, __FILE__, __LINE__);
    // device-resident keypoints come from s3d_assign_orientations: keypoint levels only, at
    // the scale of their level
    int fine = dev_kp == e->d_kp;
    for (int lv = 0; fine && lv < (int)e->g.size(); lv++) {
        const int sidx = lv % e->nlev_g + e->first_level;
        const s3d_geom &g = e->slab.empty() ? e->g[lv].g : e->slab_g[lv];
        if (sidx >= 0 && sidx < e->K && !s3d_desc_window_fine(g.scale, g.ux, g.uy, g.uz)) fine = 0;
    }
    return s3d_k_descriptors(e, dev_kp, n, (unsigned char *)dev_desc, fine);
}

int s3d_extract_descriptors(s3d_engine *e, const s3d_keypoint *kp, int n, void *host_desc)
{
    S3dRange nvtx_range("s3d:descriptors");
    DeviceGuard guard(e->device);
    if (n < 1) return s3d_fail(e, "s3d_extract_descriptors: n < 1", cudaSuccess, __FILE__, __LINE__);
    if (e->noct < 1)
        return s3d_fail(e, "s3d_extract_descriptors: no pyramid", cudaSuccess, __FILE__, __LINE__);
    int kp_levels_only = 1;  // every keypoint on a level s = 0..K-1 (those have gradient volumes)
    size_t memo_lv = (size_t)-1;  // detector keypoints of a level share sd: test each (level, sd) once
    double memo_sd = 0.0;
    for (int i = 0; i < n; i++) {
        if (kp[i].o < 0 || kp[i].o >= e->noct || kp[i].s < e->first_level ||
            kp[i].s > e->first_level + e->nlev_g - 1)
            return s3d_fail(e, "s3d_extract_descriptors: keypoint octave/level outside the pyramid",
                            cudaSuccess, __FILE__, __LINE__);
        if (kp[i].s < 0 || kp[i].s >= e->K) kp_levels_only = 0;
        if (kp_levels_only) {
            const size_t lv = (size_t)kp[i].o * e->nlev_g + (kp[i].s - e->first_level);
            if (lv != memo_lv || kp[i].sd != memo_sd) {
                const s3d_geom &g = e->slab.empty() ? e->g[lv].g : e->slab_g[lv];
                if (!s3d_desc_window_fine(kp[i].sd, g.ux, g.uy, g.uz)) kp_levels_only = 0;
                memo_lv = lv, memo_sd = kp[i].sd;
            }
        }
    }
    if (n > e->kp_in_cap) {
        if (e->d_kp_in) cudaFree(e->d_kp_in);
        e->d_kp_in = nullptr;
        e->kp_in_cap = 0;
        S3D_CUDA(e, cudaMalloc(&e->d_kp_in, (size_t)n * sizeof(s3d_keypoint)));
        e->kp_in_cap = n;
    }
    const size_t bytes = (size_t)n * S3D_DESC_STRIDE;
    if (bytes > e->desc_cap) {
        if (e->d_desc) cudaFree(e->d_desc);
        e->d_desc = nullptr;
        e->desc_cap = 0;
        S3D_CUDA(e, cudaMalloc(&e->d_desc, bytes));
        e->desc_cap = bytes;
    }
    S3D_CUDA(e, cudaMemcpyAsync(e->d_kp_in, kp, (size_t)n * sizeof(s3d_keypoint), cudaMemcpyHostToDevice, e->stream));
    // chunked: every chunk's kernel is queued first (an event after each); the D2H copies then
    // run on a second stream behind their events.  The caller's buffer is malloc memory (the
    // reference's ownership contract), and a D2H copy into pageable memory blocks
    // the host until it is done -- so the copies must be issued AFTER all launches, or chunk
    // i+1's kernel would wait for chunk i's copy.  This way only the last chunk's copy is
    // exposed; the rest leave the device behind the compute.
    // The chunks' kernels alternate between the engine's stream and a second compute stream:
    // kernels of one stream run one after the other, and every chunk would end with a tail of
    // draining CTAs (a keypoint's CTA runs 0.1 ... 0.6 ms, 592 are resident) before the next
    // chunk starts -- 15 tails per 512^3 volume.  On two streams the next chunk's CTAs fill the
    // SMs the draining chunk leaves (option "desc_streams", 1 = one stream).
    if (!e->copy_stream) S3D_CUDA(e, cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
    const int chunk = std::max(256, e->opt_desc_chunk);
    const int nchunks = (n + chunk - 1) / chunk;
    const bool two = e->opt_desc_streams >= 2 && nchunks > 1;
    cudaStream_t main_stream = e->stream;
    cudaEvent_t ready = nullptr;
    if (two) {
        if (!e->aux_stream) S3D_CUDA(e, cudaStreamCreateWithFlags(&e->aux_stream, cudaStreamNonBlocking));
        // gradient volumes (first call after a detect) and the keypoint upload happen on the
        // engine's stream: the second stream starts behind them
        if (s3d_gradients_prepare(e)) return -1;
        S3D_CUDA(e, cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
        if (cudaEventRecord(ready, main_stream) != cudaSuccess ||
            cudaStreamWaitEvent(e->aux_stream, ready, 0) != cudaSuccess) {
            cudaEventDestroy(ready);
            return s3d_fail(e, "descriptor streams", cudaGetLastError(), __FILE__, __LINE__);
        }
    }
    std::vector<cudaEvent_t> done(nchunks, nullptr);
    int rc = 0;
    for (int c = 0; c < nchunks && !rc; c++) {
        const int lo = c * chunk, cnt = std::min(chunk, n - lo);
        cudaStream_t st = two && (c & 1) ? e->aux_stream : main_stream;
        e->stream = st;
        rc = s3d_k_descriptors(e, e->d_kp_in + lo, cnt, e->d_desc + (size_t)lo * S3D_DESC_STRIDE,
                               kp_levels_only);
        e->stream = main_stream;
        if (rc) break;
        if (cudaEventCreateWithFlags(&done[c], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventRecord(done[c], st) != cudaSuccess)
            rc = s3d_fail(e, "descriptor events", cudaGetLastError(), __FILE__, __LINE__);
    }
    for (int c = 0; c < nchunks && !rc; c++) {
        const int lo = c * chunk, cnt = std::min(chunk, n - lo);
        if (cudaStreamWaitEvent(e->copy_stream, done[c], 0) != cudaSuccess ||
            cudaMemcpyAsync((unsigned char *)host_desc + (size_t)lo * S3D_DESC_STRIDE,
                            e->d_desc + (size_t)lo * S3D_DESC_STRIDE, (size_t)cnt * S3D_DESC_STRIDE,
                            cudaMemcpyDeviceToHost, e->copy_stream) != cudaSuccess)
            rc = s3d_fail(e, "descriptor download", cudaGetLastError(), __FILE__, __LINE__);
    }
    cudaError_t ce = cudaStreamSynchronize(main_stream);
    if (two) {  // later work on the engine's stream is ordered behind the second stream's chunks
        const cudaError_t ca = cudaStreamSynchronize(e->aux_stream);
        if (ce == cudaSuccess) ce = ca;
        cudaEventDestroy(ready);
    }
    cudaError_t ce2 = cudaStreamSynchronize(e->copy_stream);
    for (int c = 0; c < nchunks; c++)
        if (done[c]) cudaEventDestroy(done[c]);
    if (rc) return -1;
    if (ce != cudaSuccess || ce2 != cudaSuccess)
        return s3d_fail(e, "descriptor download", ce != cudaSuccess ? ce : ce2, __FILE__, __LINE__);
    return 0;
}

int s3d_single_level(s3d_engine *e, const float *host, int nx, int ny, int nz, size_t xs, size_t ys,
                     size_t zs, const double units[3], double scale, const s3d_filter *smooth)
{  // SIFT3D_extract_raw_descriptors / SIFT3D_assign_orientations: a one-level pyramid
   // (first_octave 0, first_level 0; sift.c:2140-2165) holding smooth_scale_raw_input
    DeviceGuard guard(e->device);
    S3D_CUDA(e, cudaStreamSynchronize(e->stream));
    TapSet ts;
    if (to_tapset(e, smooth, ts)) return -1;
    const bool same = e->noct == 1 && e->nlev_g == 1 && e->first_level == 0 && e->g.size() == 1 &&
                      e->g[0].g.nx == nx && e->g[0].g.ny == ny && e->g[0].g.nz == nz;
    if (!same) {
        free_pyramid(e);
        e->noct = 1;
        e->K = 1;
        e->nlev_g = 1;
        e->nlev_d = 0;
        e->first_level = 0;
        e->g.resize(1);
        e->g[0].g.nx = nx;
        e->g[0].g.ny = ny;
        e->g[0].g.nz = nz;
        S3D_CUDA(e, cudaMalloc(&e->g[0].d, (size_t)nx * ny * nz * sizeof(float)));
        S3D_CUDA(e, cudaMalloc(&e->d_level_ptrs, sizeof(float *)));
        S3D_CUDA(e, cudaMalloc(&e->d_level_dims, 3 * sizeof(int)));
        S3D_CUDA(e, cudaMalloc(&e->d_level_units, 3 * sizeof(float)));
        S3D_CUDA(e, cudaMalloc(&e->d_level_scales, sizeof(double)));
        e->n_scalars = 1;
        S3D_CUDA(e, cudaMalloc(&e->d_scalars, sizeof(unsigned)));
    }
    e->g[0].g.ux = units[0];
    e->g[0].g.uy = units[1];
    e->g[0].g.uz = units[2];
    e->g[0].g.scale = scale;
    const int dims[3] = {nx, ny, nz};
    const float fu[3] = {(float)units[0], (float)units[1], (float)units[2]};
    S3D_CUDA(e, cudaMemcpy(e->d_level_ptrs, &e->g[0].d, sizeof(float *), cudaMemcpyHostToDevice));
    S3D_CUDA(e, cudaMemcpy(e->d_level_dims, dims, sizeof(dims), cudaMemcpyHostToDevice));
    S3D_CUDA(e, cudaMemcpy(e->d_level_units, fu, sizeof(fu), cudaMemcpyHostToDevice));
    S3D_CUDA(e, cudaMemcpy(e->d_level_scales, &scale, sizeof(double), cudaMemcpyHostToDevice));
    e->grad_valid = false;
    if (ensure_im(e, nx, ny, nz)) return -1;
    if (upload_strided(e, e->im, host, nx, ny, nz, xs, ys, zs)) return -1;
    const float uf[3] = {(float)(1.0 / units[0]), (float)(1.0 / units[1]), (float)(1.0 / units[2])};
    if (s3d_k_blur(e, e->im, e->g[0].d, nx, ny, nz, 1, ts, uf)) return -1;
    const size_t n = (size_t)nx * ny * nz;
    if (s3d_k_max_abs(e, e->g[0].d, n, e->d_scalars)) return -1;
    if (s3d_k_scale(e, e->g[0].d, e->g[0].d, n, e->d_scalars)) return -1;
    return 0;
}

int s3d_orient_keypoints(s3d_engine *e, s3d_keypoint *kp, int n, double sig_fctr,
                         double corner_thresh, double *conf, unsigned char *ok)
{
    DeviceGuard guard(e->device);
    if (n < 1 || e->noct < 1)
        return s3d_fail(e, "s3d_orient_keypoints: no keypoints / pyramid", cudaSuccess, __FILE__, __LINE__);
    s3d_keypoint *d_kp = nullptr;
    double *d_conf = nullptr;
    unsigned char *d_ok = nullptr;
    int rc = -1;
    do {
        if (cudaMalloc(&d_kp, (size_t)n * sizeof(s3d_keypoint)) != cudaSuccess ||
            cudaMalloc(&d_conf, (size_t)n * sizeof(double)) != cudaSuccess ||
            cudaMalloc(&d_ok, (size_t)n) != cudaSuccess) {
            s3d_fail(e, "s3d_orient_keypoints: cudaMalloc", cudaGetLastError(), __FILE__, __LINE__);
            break;
        }
        cudaError_t ce = cudaMemcpyAsync(d_kp, kp, (size_t)n * sizeof(s3d_keypoint), cudaMemcpyHostToDevice, e->stream);
        if (ce != cudaSuccess) { s3d_fail(e, "upload", ce, __FILE__, __LINE__); break; }
        if (s3d_k_orient_list(e, d_kp, n, sig_fctr, corner_thresh, d_ok, d_conf)) break;
        ce = cudaMemcpyAsync(kp, d_kp, (size_t)n * sizeof(s3d_keypoint), cudaMemcpyDeviceToHost, e->stream);
        if (ce == cudaSuccess && conf)
            ce = cudaMemcpyAsync(conf, d_conf, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, e->stream);
        if (ce == cudaSuccess && ok)
            ce = cudaMemcpyAsync(ok, d_ok, (size_t)n, cudaMemcpyDeviceToHost, e->stream);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
        if (ce != cudaSuccess) { s3d_fail(e, "download", ce, __FILE__, __LINE__); break; }
        rc = 0;
    } while (0);
    cudaStreamSynchronize(e->stream);
    if (d_kp) cudaFree(d_kp);
    if (d_conf) cudaFree(d_conf);
    if (d_ok) cudaFree(d_ok);
    return rc;
}

int s3d_blur_device(s3d_engine *e, const float *dev_src, float *dev_dst, int nx, int ny, int nz,
                    int nc, const float *taps, int width, double unit, const double units[3])
{
    DeviceGuard guard(e->device);
    s3d_filter f = {taps, width};
    TapSet t;
    if (to_tapset(e, &f, t)) return -1;
    const float uf[3] = {(float)(unit / units[0]), (float)(unit / units[1]), (float)(unit / units[2])};
    return s3d_k_blur(e, dev_src, dev_dst, nx, ny, nz, nc, t, uf);
}

// wall-clock trace of the dense path (S3D_TRACE=1)
static double now_ms()
{
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return 1e3 * (double)ts.tv_sec + 1e-6 * (double)ts.tv_nsec;
}

static int dense_ensure(s3d_engine *e, int i, size_t bytes)
{
    if (bytes <= e->dense_cap[i]) return 0;
    if (e->dense_buf[i]) cudaFree(e->dense_buf[i]);
    e->dense_buf[i] = nullptr;
    e->dense_cap[i] = 0;
    S3D_CUDA(e, cudaMalloc(&e->dense_buf[i], bytes));
    e->dense_cap[i] = bytes;
    return 0;
}

static int stage_ensure(s3d_engine *e, size_t bytes)
{
    if (e->stage_cap >= bytes) return 0;
    for (int i = 0; i < 2; i++) {
        if (e->stage[i]) cudaFreeHost(e->stage[i]);
        e->stage[i] = nullptr;
    }
    e->stage_cap = 0;
    for (int i = 0; i < 2; i++) S3D_CUDA(e, cudaHostAlloc(&e->stage[i], bytes, cudaHostAllocDefault));
    e->stage_cap = bytes;
    return 0;
}

// A persistent team of host threads for the staged copies to / from pageable memory (creating
// threads per 32 MB chunk cost ~3 ms per 512 MB volume).  One team per process, created on
// first use; its size is the core count divided by the ranks sharing the host
// ($LOCAL_WORLD_SIZE, set by torchrun), at most 8 ($S3D_COPY_THREADS overrides).
static void par_memcpy(char *d, const char *src, size_t len) { HostTeam::get().copy(d, src, len); }

// Pinned ring + events of the chunk pipeline (kept between calls).
static int pipe_ensure(s3d_engine *e, size_t bytes, size_t ns)
{
    if (e->pipe_cap < bytes) {
        if (e->pipe_buf) cudaFreeHost(e->pipe_buf);
        e->pipe_buf = nullptr;
        e->pipe_cap = 0;
        S3D_CUDA(e, cudaHostAlloc(&e->pipe_buf, bytes, cudaHostAllocDefault));
        e->pipe_cap = bytes;
    }
    while (e->pipe_ev.size() < ns) {
        cudaEvent_t ev = nullptr;
        S3D_CUDA(e, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        e->pipe_ev.push_back(ev);
    }
    return 0;
}

// One transfer through the chunk pipeline (see PipeJob): dir 0 = pageable host -> device,
// dir 1 = device -> pageable host.  Returns with the transfer complete.
static int pipe_transfer(s3d_engine *e, int dir, void *dev, void *host, size_t bytes)
{
    // Ring geometry: 16 slots of 2 MB measured best on one GPU (tools/copy_pipe_ab.py); with
    // several ranks on the host the ring shrinks to 8 x 1 MB so that the rings of all ranks stay
    // inside the last-level cache together, and the polling caller replaces one worker.
    HostTeam &team = HostTeam::get();
    const bool shared_host = team.ranks_on_host() > 1;
    const size_t ch = (size_t)(e->opt_pipe_chunk_kb > 0 ? std::max(256, e->opt_pipe_chunk_kb) : (shared_host ? 1024 : 2048)) << 10;
    const size_t ns = (size_t)(e->opt_pipe_slots > 0 ? std::min(64, std::max(2, e->opt_pipe_slots)) : (shared_host ? 8 : 16));
    if (pipe_ensure(e, ch * ns, ns)) return -1;
    PipeJob J;
    J.dir = dir;
    J.host = (char *)host;
    J.slots = (char *)e->pipe_buf;
    J.bytes = bytes, J.ch = ch, J.ns = ns;
    J.nch = (bytes + ch - 1) / ch;
    std::unique_ptr<std::atomic<unsigned char>[]> flags(new std::atomic<unsigned char>[J.nch]);
    for (size_t c = 0; c < J.nch; c++) flags[c].store(0, std::memory_order_relaxed);
    J.flag = flags.get();
    cudaError_t ce = cudaSuccess;
    cudaStream_t st = e->stream;
    // issue the DMA of chunk c (in order, on the engine's stream) / poll the event of a slot
    auto issue = [&](size_t c) {
        char *slot = J.slots + (c % ns) * ch;
        char *d = (char *)dev + c * ch;
        const size_t len = std::min(ch, bytes - c * ch);
        ce = dir == 0 ? cudaMemcpyAsync(d, slot, len, cudaMemcpyHostToDevice, st)
                      : cudaMemcpyAsync(slot, d, len, cudaMemcpyDeviceToHost, st);
        if (ce == cudaSuccess) ce = cudaEventRecord(e->pipe_ev[c % ns], st);
        return ce == cudaSuccess;
    };
    auto poll = [&](size_t c) {  // 1 done, 0 in flight, -1 error
        const cudaError_t q = cudaEventQuery(e->pipe_ev[c % ns]);
        if (q == cudaSuccess) return 1;
        if (q == cudaErrorNotReady) return 0;
        ce = q;
        return -1;
    };
    s3d_pipe_run(J, issue, poll, shared_host && team.workers() > 1 ? 1u : 0u);
    const cudaError_t cs = cudaStreamSynchronize(st);  // the last DMAs; the ring is free again
    cudaGetLastError();  // cudaErrorNotReady of the polls is not an error
    if (ce == cudaSuccess) ce = cs;
    if (ce != cudaSuccess) return s3d_fail(e, dir ? "pipelined download" : "pipelined upload", ce, __FILE__, __LINE__);
    return 0;
}

// The pipeline without a polling thread (host_pipe.h: Pipe2Job; option copy_pipe = 2): every
// participant has its own stream, two pinned slots and two events.
static int pipe2_transfer(s3d_engine *e, int dir, void *dev, void *host, size_t bytes)
{
    HostTeam &team = HostTeam::get();
    const unsigned P = team.workers() + 1;
    const size_t ch = (size_t)(e->opt_pipe_chunk_kb > 0 ? std::max(256, e->opt_pipe_chunk_kb)
                                                         : (team.ranks_on_host() > 1 ? 1024 : 2048)) << 10;
    if (pipe_ensure(e, ch * 2 * P, 2 * P)) return -1;
    while (e->pipe_streams.size() < P) {
        cudaStream_t st = nullptr;
        S3D_CUDA(e, cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        e->pipe_streams.push_back(st);
    }
    // the device buffer is produced / last read on the engine's stream
    S3D_CUDA(e, cudaStreamSynchronize(e->stream));
    Pipe2Job J;
    J.dir = dir;
    J.host = (char *)host;
    J.slots = (char *)e->pipe_buf;
    J.bytes = bytes, J.ch = ch;
    J.nch = (bytes + ch - 1) / ch;
    std::atomic<int> first_err{0};
    auto note = [&](cudaError_t ce) {
        int zero = 0;
        if (ce != cudaSuccess) first_err.compare_exchange_strong(zero, (int)ce);
        return ce == cudaSuccess;
    };
    const int device = e->device;
    auto enter = [&](unsigned) { return note(cudaSetDevice(device)); };
    auto issue = [&](unsigned p, size_t c, int k) {
        char *slot = J.slots + (size_t)(2 * p + k) * ch;
        char *d = (char *)dev + c * ch;
        const size_t len = std::min(ch, bytes - c * ch);
        cudaStream_t st = e->pipe_streams[p];
        if (!note(dir == 0 ? cudaMemcpyAsync(d, slot, len, cudaMemcpyHostToDevice, st)
                           : cudaMemcpyAsync(slot, d, len, cudaMemcpyDeviceToHost, st)))
            return false;
        return note(cudaEventRecord(e->pipe_ev[2 * p + k], st));
    };
    auto wait = [&](unsigned p, int k) { return note(cudaEventSynchronize(e->pipe_ev[2 * p + k])); };
    // ranks sharing the host: the caller takes worker 0's place, so that a rank runs exactly the
    // cores / ranks threads the team was sized for
    const bool ok = s3d_pipe2_run(J, enter, issue, wait, team.ranks_on_host() > 1 && team.workers() > 1 ? 1u : 0u);
    cudaError_t ce = ok ? cudaSuccess : (cudaError_t)first_err.load();
    for (unsigned p = 0; p < P; p++) {  // join: the last DMAs of every participant
        const cudaError_t cs = cudaStreamSynchronize(e->pipe_streams[p]);
        if (ce == cudaSuccess) ce = cs;
    }
    if (!ok && ce == cudaSuccess) ce = cudaErrorUnknown;
    if (ce != cudaSuccess) return s3d_fail(e, dir ? "pipelined download" : "pipelined upload", ce, __FILE__, __LINE__);
    return 0;
}

// Which staged path a >= 32 MB copy takes (option "copy_pipe": 0 / 1 / 2 force one, -1 = automatic).
// Measured on a 16-core host (profiles/r02_copy_pipe_ab.txt).  With the 8 threads a process gets
// when it has the host to itself, the pipeline with a polling caller (1) and the poller-free one
// (2) are equal (512^3 detect call 27.5 / 27.9 ms, dense 256^3 call 28.8 / 28.1 ms) and both beat
// the chunk-at-a-time copy on the dense download (30.7 ms).  With the 4 / 2 threads a rank gets
// when 4 / 8 ranks share that host ($LOCAL_WORLD_SIZE) a thread that only polls is a quarter /
// half of the copy capacity: detect 37.8 / 73.4 ms for (1) against 33.3 / 44.9 ms for the
// chunk-at-a-time copy (0), where the caller copies too -- and 31.3 / 36.7 ms against 42.0 /
// 40.5 ms (same run) for (2), where everybody copies and nobody meets anybody.  So: (1) alone on
// the host, (2) when ranks share it.
static int copy_pipe_mode(const s3d_engine *e)
{
    if (e->opt_copy_pipe >= 0) return e->opt_copy_pipe;
    return HostTeam::get().ranks_on_host() == 1 ? 1 : 2;
}

// Device -> PAGEABLE host memory (the caller's malloc'ed Image, SURVEY.md 8b ownership rule).
// A plain cudaMemcpy stages through the driver's bounce buffer and copies out on one core; here
// the DMA lands in a ring of two pinned buffers and a team of host threads copies each chunk
// out (and takes the first-touch page faults of a fresh allocation in parallel) while the DMA
// fills the other buffer.
static int d2h_pageable(s3d_engine *e, void *dst, const void *dev, size_t bytes)
{
    const size_t CH = (size_t)32 << 20;
    if (!e->opt_dense_copy || bytes < 2 * CH) {
        S3D_CUDA(e, cudaMemcpyAsync(dst, dev, bytes, cudaMemcpyDeviceToHost, e->stream));
        S3D_CUDA(e, cudaStreamSynchronize(e->stream));
        return 0;
    }
    if (copy_pipe_mode(e) == 2) return pipe2_transfer(e, 1, const_cast<void *>(dev), dst, bytes);
    if (copy_pipe_mode(e) == 1) return pipe_transfer(e, 1, const_cast<void *>(dev), dst, bytes);
    if (stage_ensure(e, CH)) return -1;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    for (int i = 0; i < 2; i++) S3D_CUDA(e, cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
    const size_t nch = (bytes + CH - 1) / CH;
    cudaError_t ce = cudaSuccess;
    auto issue = [&](size_t c) {
        const size_t off = c * CH, len = std::min(CH, bytes - off);
        ce = cudaMemcpyAsync(e->stage[c & 1], (const char *)dev + off, len, cudaMemcpyDeviceToHost, e->stream);
        if (ce == cudaSuccess) ce = cudaEventRecord(ev[c & 1], e->stream);
    };
    issue(0);
    for (size_t c = 0; c < nch && ce == cudaSuccess; c++) {
        ce = cudaEventSynchronize(ev[c & 1]);
        if (ce != cudaSuccess) break;
        if (c + 1 < nch) issue(c + 1);  // the other buffer: free since chunk c-1 was copied out
        const size_t off = c * CH, len = std::min(CH, bytes - off);
        par_memcpy((char *)dst + off, (const char *)e->stage[c & 1], len);
    }
    cudaStreamSynchronize(e->stream);
    for (int i = 0; i < 2; i++) cudaEventDestroy(ev[i]);
    if (ce != cudaSuccess) return s3d_fail(e, "staged download", ce, __FILE__, __LINE__);
    return 0;
}

// PAGEABLE host memory -> device, the mirror image: a team of host threads fills one pinned
// buffer while the DMA drains the other.
static int h2d_pageable(s3d_engine *e, void *dev, const void *host, size_t bytes)
{
    const size_t CH = (size_t)16 << 20;
    if (!e->opt_dense_copy || bytes < 2 * CH) {
        S3D_CUDA(e, cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, e->stream));
        return 0;
    }
    if (copy_pipe_mode(e) == 2) return pipe2_transfer(e, 0, dev, const_cast<void *>(host), bytes);
    if (copy_pipe_mode(e) == 1) return pipe_transfer(e, 0, dev, const_cast<void *>(host), bytes);
    if (stage_ensure(e, CH)) return -1;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    for (int i = 0; i < 2; i++) S3D_CUDA(e, cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
    const size_t nch = (bytes + CH - 1) / CH;
    cudaError_t ce = cudaStreamSynchronize(e->stream);  // earlier users of the staging ring
    for (size_t c = 0; c < nch && ce == cudaSuccess; c++) {
        if (c >= 2) ce = cudaEventSynchronize(ev[c & 1]);  // its previous DMA has drained
        if (ce != cudaSuccess) break;
        const size_t off = c * CH, len = std::min(CH, bytes - off);
        par_memcpy((char *)e->stage[c & 1], (const char *)host + off, len);
        ce = cudaMemcpyAsync((char *)dev + off, e->stage[c & 1], len, cudaMemcpyHostToDevice, e->stream);
        if (ce == cudaSuccess) ce = cudaEventRecord(ev[c & 1], e->stream);
    }
    cudaStreamSynchronize(e->stream);
    for (int i = 0; i < 2; i++) cudaEventDestroy(ev[i]);
    if (ce != cudaSuccess) return s3d_fail(e, "staged upload", ce, __FILE__, __LINE__);
    return 0;
}

int s3d_dense_descriptors(s3d_engine *e, const float *host_in, int nx, int ny, int nz, size_t xs,
                          size_t ys, size_t zs, const double units[3],
                          const double desc_units[3], const s3d_filter *smooth,
                          const s3d_filter *window, float *host_out)
{
    S3dRange nvtx_range("s3d:dense");
    DeviceGuard guard(e->device);
    const size_t n = (size_t)nx * ny * nz;
    TapSet ts, tw;
    if (to_tapset(e, smooth, ts) || to_tapset(e, window, tw)) return -1;
    const bool trace = getenv("S3D_TRACE") != nullptr;
    double t[8];
    t[0] = now_ms();
    if (dense_ensure(e, 0, n * 4) || dense_ensure(e, 1, n * 4) || dense_ensure(e, 2, n * 48) ||
        dense_ensure(e, 3, n * 48))
        return -1;
    float *raw = e->dense_buf[0], *sm = e->dense_buf[1], *t12 = e->dense_buf[2], *d12 = e->dense_buf[3];
    int rc = -1;
    for (int i = 0; i < 4; i++)  // device-side stage times of the call (s3d_dense_last_timing)
        if (!e->dense_ev[i] && cudaEventCreate(&e->dense_ev[i]) != cudaSuccess) e->dense_ev[i] = nullptr;
    const bool have_ev = e->dense_ev[0] && e->dense_ev[1] && e->dense_ev[2] && e->dense_ev[3];
    e->dense_ms[0] = e->dense_ms[1] = e->dense_ms[2] = -1.0;
    do {
        if (have_ev) cudaEventRecord(e->dense_ev[0], e->stream);
        if (upload_strided(e, raw, host_in, nx, ny, nz, xs, ys, zs)) break;
        if (have_ev) cudaEventRecord(e->dense_ev[1], e->stream);
        if (trace) cudaStreamSynchronize(e->stream);
        t[1] = now_ms();
        // smooth_scale_raw_input (sift.c:1978-2006)
        const float uf[3] = {(float)(1.0 / units[0]), (float)(1.0 / units[1]), (float)(1.0 / units[2])};
        if (s3d_k_blur(e, raw, sm, nx, ny, nz, 1, ts, uf)) break;
        unsigned *mx = e->d_counter ? reinterpret_cast<unsigned *>(e->d_counter + 2) : nullptr;
        if (s3d_k_max_abs(e, sm, n, mx)) break;
        if (s3d_k_scale(e, sm, sm, n, mx)) break;
        // gradient direction -> barycentric channel image (sift.c:2462-2480)
        const float fu[3] = {(float)units[0], (float)units[1], (float)units[2]};
        const float iu[3] = {1.0f / fu[0], 1.0f / fu[1], 1.0f / fu[2]};
        if (s3d_k_dense(e, sm, raw, nx, ny, nz, iu, t12)) break;
        if (trace) cudaStreamSynchronize(e->stream);
        t[2] = now_ms();
        // 12-channel window blur in the units of the caller's desc image (sift.c:2451, 2483)
        const float ufd[3] = {(float)(1.0 / desc_units[0]), (float)(1.0 / desc_units[1]),
                              (float)(1.0 / desc_units[2])};
        if (s3d_k_blur(e, t12, d12, nx, ny, nz, 12, tw, ufd)) break;
        if (trace) cudaStreamSynchronize(e->stream);
        t[3] = now_ms();
        if (s3d_k_dense_post(e, d12, raw, n)) break;
        if (have_ev) cudaEventRecord(e->dense_ev[2], e->stream);
        if (trace) cudaStreamSynchronize(e->stream);
        t[4] = now_ms();
        if (d2h_pageable(e, host_out, d12, n * 48)) break;
        if (have_ev) cudaEventRecord(e->dense_ev[3], e->stream);
        t[5] = now_ms();
        rc = 0;
    } while (0);
    cudaStreamSynchronize(e->stream);
    if (rc == 0 && have_ev)
        for (int i = 0; i < 3; i++) {
            float ms = -1.0f;
            if (cudaEventElapsedTime(&ms, e->dense_ev[i], e->dense_ev[i + 1]) == cudaSuccess) e->dense_ms[i] = ms;
        }
    if (trace && rc == 0)
        fprintf(stderr, "[s3d dense %dx%dx%d] alloc %.2f upload %.2f smooth+bary %.2f blur12 %.2f post %.2f "
                        "download %.2f ms\n", nx, ny, nz, 0.0, t[1] - t[0], t[2] - t[1], t[3] - t[2],
                t[4] - t[3], t[5] - t[4]);
    return rc;
}

int s3d_dense_last_timing(const s3d_engine *e, double ms[3])
{
    for (int i = 0; i < 3; i++) ms[i] = e->dense_ms[i];
    return e->dense_ms[1] >= 0.0 ? 0 : -1;
}

int s3d_dense_descriptors_rotate(s3d_engine *e, const float *host_in, int nx, int ny, int nz,
                                 size_t xs, size_t ys, size_t zs, const double units[3],
                                 const s3d_filter *smooth, double ori_sigma, double desc_sigma,
                                 double corner_thresh, float *host_out)
{
    S3dRange nvtx_range("s3d:dense_rotate");
    DeviceGuard guard(e->device);
    const size_t n = (size_t)nx * ny * nz;
    TapSet ts;
    if (to_tapset(e, smooth, ts)) return -1;
    if (dense_ensure(e, 0, n * 4) || dense_ensure(e, 1, n * 4) || dense_ensure(e, 3, n * 48)) return -1;
    float *raw = e->dense_buf[0], *sm = e->dense_buf[1], *d12 = e->dense_buf[3];
    int rc = -1;
    do {
        if (upload_strided(e, raw, host_in, nx, ny, nz, xs, ys, zs)) break;
        // smooth_scale_raw_input (sift.c:1978-2006)
        const float uf[3] = {(float)(1.0 / units[0]), (float)(1.0 / units[1]), (float)(1.0 / units[2])};
        if (s3d_k_blur(e, raw, sm, nx, ny, nz, 1, ts, uf)) break;
        unsigned *mx = e->d_counter ? reinterpret_cast<unsigned *>(e->d_counter + 2) : nullptr;
        if (s3d_k_max_abs(e, sm, n, mx)) break;
        if (s3d_k_scale(e, sm, sm, n, mx)) break;
        // per-voxel orientation + rotated histogram (sift.c:2521-2588)
        const float fu[3] = {(float)units[0], (float)units[1], (float)units[2]};
        if (s3d_k_dense_rotate(e, sm, nx, ny, nz, fu, ori_sigma, desc_sigma, corner_thresh, d12))
            break;
        if (s3d_k_dense_post(e, d12, raw, n)) break;
        if (d2h_pageable(e, host_out, d12, n * 48)) break;
        rc = 0;
    } while (0);
    cudaStreamSynchronize(e->stream);
    return rc;
}

int s3d_level_download(s3d_engine *e, int which, int o, int s, float *host_dst)
{
    DeviceGuard guard(e->device);
    const int nl = which ? e->nlev_d : e->nlev_g;
    if (o < 0 || o >= e->noct || s < -1 || s > nl - 2)
        return s3d_fail(e, "s3d_level_download: index", cudaSuccess, __FILE__, __LINE__);
    const LevelDev &l = which ? e->dog[(size_t)o * nl + s + 1] : e->g[(size_t)o * nl + s + 1];
    S3D_CUDA(e, cudaMemcpyAsync(host_dst, l.d, l.n() * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    S3D_CUDA(e, cudaStreamSynchronize(e->stream));
    return 0;
}

int s3d_pyramid_copy(s3d_engine *dst, const s3d_engine *src)
{
    if (src->noct == 0) return 0;
    std::vector<s3d_geom> g(src->g.size()), d(src->dog.size());
    for (size_t i = 0; i < g.size(); i++) g[i] = src->g[i].g;
    for (size_t i = 0; i < d.size(); i++) d[i] = src->dog[i].g;
    if (s3d_pyramid_resize(dst, src->noct, src->K, g.data(), d.data())) return -1;
    DeviceGuard guard(dst->device);
    dst->grad_valid = false;
    cudaStreamSynchronize(src->stream);
    for (size_t i = 0; i < g.size(); i++)
        S3D_CUDA(dst, cudaMemcpyPeerAsync(dst->g[i].d, dst->device, src->g[i].d, src->device,
                                          src->g[i].n() * sizeof(float), dst->stream));
    for (size_t i = 0; i < d.size(); i++)
        S3D_CUDA(dst, cudaMemcpyPeerAsync(dst->dog[i].d, dst->device, src->dog[i].d, src->device,
                                          src->dog[i].n() * sizeof(float), dst->stream));
    dst->first_taps = src->first_taps;
    dst->oct_taps = src->oct_taps;
    S3D_CUDA(dst, cudaStreamSynchronize(dst->stream));
    return 0;
}

int s3d_debug_read(s3d_engine *e, void *host, size_t bytes)
{
    DeviceGuard guard(e->device);
    if (!e->d_blur_dbg) return -1;
    S3D_CUDA(e, cudaStreamSynchronize(e->stream));
    S3D_CUDA(e, cudaMemcpy(host, e->d_blur_dbg, bytes, cudaMemcpyDeviceToHost));
    return 0;
}

void *s3d_dev_alloc(s3d_engine *e, size_t bytes)
{
    DeviceGuard guard(e->device);
    void *p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
        s3d_fail(e, "s3d_dev_alloc", cudaGetLastError(), __FILE__, __LINE__);
        return nullptr;
    }
    return p;
}

void s3d_dev_free(s3d_engine *e, void *p)
{
    DeviceGuard guard(e->device);
    if (p) cudaFree(p);
}

int s3d_memcpy_h2d(s3d_engine *e, void *dev, const void *host, size_t bytes)
{
    DeviceGuard guard(e->device);
    S3D_CUDA(e, cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, e->stream));
    S3D_CUDA(e, cudaStreamSynchronize(e->stream));
    return 0;
}

int s3d_memcpy_d2h(s3d_engine *e, void *host, const void *dev, size_t bytes)
{
    DeviceGuard guard(e->device);
    S3D_CUDA(e, cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, e->stream));
    S3D_CUDA(e, cudaStreamSynchronize(e->stream));
    return 0;
}

int s3d_host_roundtrip(s3d_engine *e, const void *host_src, void *host_dst, size_t bytes)
{
    DeviceGuard guard(e->device);
    void *d = nullptr;
    S3D_CUDA(e, cudaMalloc(&d, bytes));
    int rc = h2d_pageable(e, d, host_src, bytes);
    if (!rc) rc = d2h_pageable(e, host_dst, d, bytes);
    cudaStreamSynchronize(e->stream);
    cudaFree(d);
    return rc;
}

}  // extern "C"
