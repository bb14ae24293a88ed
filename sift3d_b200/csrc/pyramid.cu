// pyramid.cu -- scale-space pyramid kernels for sm_100a.
//
// Replaces (reference bbrister/SIFT3D v1.4.6):
//   im_max_abs / im_scale        imutil.c:1959-1991
//   convolve_sep_gen             imutil.c:2274-2393  (via apply_Sep_FIR_filter :3459-3544)
//   im_downsample_2x             imutil.c:1742-1768
//   im_subtract (build_dog)      imutil.c:1997-2017, sift.c:1052-1071
//   detect_extrema               sift.c:1074-1212
//
// Arithmetic contract: every tap is
//     acc = acc (+) tap (*) ( (1-frac) (*) lo (+) frac (*) hi )
// with one IEEE rounding per (*) and (+), taps visited d = -hw..hw.  No FMA.
#include "common.cuh"

#include <algorithm>
#include <cfloat>
#include <cmath>

namespace {

__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---------------------------------------------------------------- max |x|
// The flat kernels below take `head` = number of leading elements before the first 16-byte
// boundary (plane ranges of a Z-slab start at arbitrary element offsets): scalar head,
// float4 body, scalar tail.
__global__ void __launch_bounds__(256) k_max_abs(const float *__restrict__ x, size_t n,
                                                 size_t head, unsigned *__restrict__ out_bits)
{
    float m = 0.0f;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nth = (size_t)gridDim.x * blockDim.x;
    const size_t n4 = (n - head) / 4;
    const float4 *x4 = reinterpret_cast<const float4 *>(x + head);
    for (size_t i = tid; i < n4; i += nth) {
        const float4 v = __ldg(x4 + i);
        m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
    for (size_t i = tid; i < head; i += nth) m = fmaxf(m, fabsf(x[i]));
    for (size_t i = head + n4 * 4 + tid; i < n; i += nth) m = fmaxf(m, fabsf(x[i]));
    m = warp_max(m);
    __shared__ float sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.0f;
        m = warp_max(m);
        // non-negative floats order like their bit patterns
        if (threadIdx.x == 0) atomicMax(out_bits, __float_as_uint(m));
    }
}

// ---------------------------------------------------------------- x / max
__global__ void __launch_bounds__(256) k_scale(const float *__restrict__ src,
                                               float *__restrict__ dst, size_t n, size_t head,
                                               const unsigned *__restrict__ max_bits)
{
    const float mx = __uint_as_float(*max_bits);
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nth = (size_t)gridDim.x * blockDim.x;
    const size_t n4 = (n - head) / 4;
    const float4 *s4 = reinterpret_cast<const float4 *>(src + head);
    float4 *d4 = reinterpret_cast<float4 *>(dst + head);
    if (mx == 0.0f) {  // im_scale returns early (imutil.c:1984-1985)
        if (src == dst) return;
        for (size_t i = tid; i < n4; i += nth) d4[i] = s4[i];
        for (size_t i = tid; i < head; i += nth) dst[i] = src[i];
        for (size_t i = head + n4 * 4 + tid; i < n; i += nth) dst[i] = src[i];
        return;
    }
    for (size_t i = tid; i < head; i += nth) dst[i] = __fdiv_rn(src[i], mx);
    for (size_t i = tid; i < n4; i += nth) {
        float4 v = s4[i];
        v.x = __fdiv_rn(v.x, mx);
        v.y = __fdiv_rn(v.y, mx);
        v.z = __fdiv_rn(v.z, mx);
        v.w = __fdiv_rn(v.w, mx);
        d4[i] = v;
    }
    for (size_t i = head + n4 * 4 + tid; i < n; i += nth) dst[i] = __fdiv_rn(src[i], mx);
}

// ---------------------------------------------------------------- generic 1-axis FIR
// Literal restatement of convolve_sep_gen for any axis, channel count and tap
// spacing `uf` (= unit / units[axis], not necessarily dyadic).  This is the
// reference-semantics path: slow but exact for every input; the fused kernel in
// blur_fused.cu is the fast path for the common case.
// `gbase`/`nbuf`: the line may be a window [gbase, gbase + nbuf) of a longer (global) line
// (Z-slab tiling): coordinates are GLOBAL -- the f32 rounding of c, of the mirrored c and of
// frac depends on the magnitude of the index -- and only the addressing is shifted.
__device__ __forceinline__ float samp_acc(float acc, float tap, float c, const float *line,
                                          size_t st, int dim_end, int gbase, int nbuf)
{
    int lo = __float2int_rz(c);
    const float frac = __fsub_rn(c, (float)lo);
    int hi = lo + 1;
    lo = min(max(lo, 0), dim_end);  // reads the reference leaves undefined are clamped
    hi = min(max(hi, 0), dim_end);
    lo = min(max(lo - gbase, 0), nbuf - 1);
    hi = min(max(hi - gbase, 0), nbuf - 1);
    const float v = __fadd_rn(__fmul_rn(__fsub_rn(1.0f, frac), __ldg(line + (size_t)lo * st)),
                              __fmul_rn(frac, __ldg(line + (size_t)hi * st)));
    return __fadd_rn(acc, __fmul_rn(tap, v));
}

// For AXIS == 2 the launch may cover only the output planes [zoff, zoff + nz) of a buffer of
// nbuf planes which is itself the window [gbase, gbase + nbuf) of a line of nglob planes:
// `src` is the base of the buffer, `dst` points at its plane zoff.
template <int AXIS>
__global__ void __launch_bounds__(256) k_conv_axis(const float *__restrict__ src,
                                                   float *__restrict__ dst, int nx, int ny,
                                                   int nz, int nc, const TapSet taps, float uf,
                                                   int zoff, int nbuf, int gbase, int nglob,
                                                   int dyadic)
{
    const size_t total = (size_t)nx * ny * nz * nc;
    const int n = AXIS == 0 ? nx : (AXIS == 1 ? ny : nglob);
    if (AXIS != 2) nbuf = n, gbase = 0;
    // Dyadic tap spacing (uf = 2^-k, every pyramid octave with power-of-two units): for an
    // interior voxel i the sample coordinate c = i - d*uf is exact in f32, so (int)c = i +
    // floor(-d*uf) and frac = c - (int)c = frac(-d*uf) do not depend on i -- one table entry
    // per tap replaces the per-tap coordinate arithmetic, with bit-identical results.
    __shared__ int s_off[S3D_MAX_TAPS];
    __shared__ float s_f[S3D_MAX_TAPS], s_omf[S3D_MAX_TAPS];
    if (dyadic) {
        for (int t = threadIdx.x; t < taps.width; t += blockDim.x) {
            const float rel = -__fmul_rn((float)(t - taps.width / 2), uf);
            const float fl = floorf(rel);
            const float f = __fsub_rn(rel, fl);
            s_off[t] = (int)fl;
            s_f[t] = f;
            s_omf[t] = __fsub_rn(1.0f, f);
        }
        __syncthreads();
    }
    const size_t st = AXIS == 0 ? (size_t)nc : (AXIS == 1 ? (size_t)nc * nx : (size_t)nc * nx * ny);
    const int hw = taps.width / 2;
    const int dim_end = n - 1;
    const int uhw = (int)ceilf(__fmul_rn((float)hw, uf));
    const int start = uhw, end = n - 1 - (uhw + 1);
    const float conv_eps = 0.1f;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        size_t r = idx / nc;
        const int x = (int)(r % nx);
        r /= nx;
        const int y = (int)(r % ny);
        const int z = (int)(r / ny);
        const int i = AXIS == 0 ? x : (AXIS == 1 ? y : z + zoff + gbase);  // global index
        const float *line = src + (idx - (size_t)(AXIS == 2 ? z : i) * st);
        float acc = 0.0f;
        if (dyadic && i >= start && i <= end) {
            const float *base = line + (size_t)(i - gbase) * st;
            for (int t = 0; t < taps.width; t++) {
                const float *p = base + (ptrdiff_t)s_off[t] * (ptrdiff_t)st;
                const float v = __fadd_rn(__fmul_rn(s_omf[t], __ldg(p)), __fmul_rn(s_f[t], __ldg(p + st)));
                acc = __fadd_rn(acc, __fmul_rn(taps.t[t], v));
            }
        } else if (i >= start && i <= end) {
            float c = (float)i;  // carried across taps (imutil.c:2335-2350)
            for (int d = -hw; d <= hw; d++) {
                const float step = __fmul_rn((float)d, uf);
                c = __fsub_rn(c, step);
                acc = samp_acc(acc, taps.t[d + hw], c, line, st, dim_end, gbase, nbuf);
                c = __fadd_rn(c, step);
            }
        } else {
            for (int d = -hw; d <= hw; d++) {
                const float step = __fmul_rn((float)d, uf);
                float c = __fsub_rn((float)i, step);
                if (__float2int_rz(c) < 0)
                    c = -c;
                else if (__float2int_rz(c) >= dim_end)
                    c = __fsub_rn(__fsub_rn(__fmul_rn(2.0f, (float)dim_end), c), conv_eps);
                acc = samp_acc(acc, taps.t[d + hw], c, line, st, dim_end, gbase, nbuf);
            }
        }
        dst[idx] = acc;
    }
}

// ---------------------------------------------------------------- 2x decimation
__global__ void __launch_bounds__(256) k_decimate(const float *__restrict__ src, int sx, int sy,
                                                  float *__restrict__ dst, int dx, int dy, int dz)
{
    const size_t total = (size_t)dx * dy * dz;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(idx % dx);
        const size_t r = idx / dx;
        const int y = (int)(r % dy);
        const int z = (int)(r / dy);
        dst[idx] = __ldg(src + (size_t)2 * x + (size_t)sx * ((size_t)2 * y + (size_t)sy * 2 * z));
    }
}

// ---------------------------------------------------------------- DoG + max|DoG|
__global__ void __launch_bounds__(256) k_dog(const float *__restrict__ a,
                                             const float *__restrict__ b, float *__restrict__ d,
                                             size_t n, size_t head,
                                             unsigned *__restrict__ max_bits)
{
    float m = 0.0f;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nth = (size_t)gridDim.x * blockDim.x;
    const size_t n4 = (n - head) / 4;
    const float4 *a4 = reinterpret_cast<const float4 *>(a + head);
    const float4 *b4 = reinterpret_cast<const float4 *>(b + head);
    float4 *d4 = reinterpret_cast<float4 *>(d + head);
    for (size_t i = tid; i < head; i += nth) {
        const float v = __fsub_rn(a[i], b[i]);
        d[i] = v;
        m = fmaxf(m, fabsf(v));
    }
    for (size_t i = tid; i < n4; i += nth) {
        const float4 va = __ldg(a4 + i), vb = __ldg(b4 + i);
        float4 v;
        v.x = __fsub_rn(va.x, vb.x);
        v.y = __fsub_rn(va.y, vb.y);
        v.z = __fsub_rn(va.z, vb.z);
        v.w = __fsub_rn(va.w, vb.w);
        d4[i] = v;
        m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
    for (size_t i = head + n4 * 4 + tid; i < n; i += nth) {
        const float v = __fsub_rn(a[i], b[i]);
        d[i] = v;
        m = fmaxf(m, fabsf(v));
    }
    m = warp_max(m);
    __shared__ float sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.0f;
        m = warp_max(m);
        if (threadIdx.x == 0) atomicMax(max_bits, __float_as_uint(m));
    }
}

// All DoG levels of one octave in one pass: every Gaussian level is read ONCE (the per-level
// kernel reads each inner level twice): (NL + NL-1) x 4 B/voxel instead of (NL-1) x 12.
#define DOG_MAX_LEVELS 18
struct DogLevels {
    const float *g[DOG_MAX_LEVELS];
    float *d[DOG_MAX_LEVELS - 1];
    unsigned *maxbits;  // [NL-1] consecutive slots
    int nl;
};

template <int NLT>  // NLT > 0: number of Gaussian levels known at compile time
__global__ void __launch_bounds__(256) k_dog_octave(const DogLevels L, size_t n4)
{   // n4 float4 groups (n % 4 == 0, 16-byte aligned levels: the launcher checks)
    const int nl = NLT > 0 ? NLT : L.nl;
    constexpr int NA = NLT > 0 ? NLT : DOG_MAX_LEVELS;
    float m[NA - 1];
#pragma unroll
    for (int s = 0; s < NA - 1; s++) m[s] = 0.0f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
         i += (size_t)gridDim.x * blockDim.x) {
        float4 v[NA];
#pragma unroll
        for (int s = 0; s < NA; s++)
            if (s < nl) v[s] = __ldg(reinterpret_cast<const float4 *>(L.g[s]) + i);
#pragma unroll
        for (int s = 0; s < NA - 1; s++) {
            if (s >= nl - 1) break;
            float4 d;
            d.x = __fsub_rn(v[s].x, v[s + 1].x);
            d.y = __fsub_rn(v[s].y, v[s + 1].y);
            d.z = __fsub_rn(v[s].z, v[s + 1].z);
            d.w = __fsub_rn(v[s].w, v[s + 1].w);
            reinterpret_cast<float4 *>(L.d[s])[i] = d;
            m[s] = fmaxf(m[s], fmaxf(fmaxf(fabsf(d.x), fabsf(d.y)), fmaxf(fabsf(d.z), fabsf(d.w))));
        }
    }
    __shared__ float sm[8];
#pragma unroll
    for (int s = 0; s < NA - 1; s++) {
        if (s >= nl - 1) break;
        const float w = warp_max(m[s]);
        if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = w;
        __syncthreads();
        if (threadIdx.x < 32) {
            float t = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.0f;
            t = warp_max(t);
            if (threadIdx.x == 0) atomicMax(L.maxbits + s, __float_as_uint(t));
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------- extrema
// Pass A: every block scans EXT_CHUNK consecutive voxels of the octave (linear index = scan
// order), all K keypoint levels; a warp handles 32 consecutive voxels per step and lane 0 writes
// their hit mask.  Emits the bit masks and per-block counts per level, so that pass C can compact
// in the reference's (o, s, z, y, x) order.  The centre values of EXT_UNROLL steps x K levels are
// loaded before any test (independent loads in flight; the tests themselves are rare-path).
#define EXT_BLOCK 256
#define EXT_CHUNK 8192  // voxels per block = 256 mask words
#define EXT_UNROLL 4
#define EXT_MAX_LEVELS 16

struct ExtLevels {
    const float *dog[EXT_MAX_LEVELS + 2];  // dog[s+1], s = -1..K
    const unsigned *maxbits;               // dogmax bits for this octave: [s+1]
    int K;
};

template <int KT>  // KT > 0: K == KT known at compile time; KT == 0: any K <= EXT_MAX_LEVELS
__global__ void __launch_bounds__(EXT_BLOCK, KT > 0 ? 3 : 1)
    k_extrema_mark(const ExtLevels L, int nx, int ny, int nz, double peak_thresh,
                   unsigned *__restrict__ mask, int *__restrict__ blockcnt, int nblocks,
                   size_t words_per_level)
{
    const int K = KT > 0 ? KT : L.K;
    constexpr int KA = KT > 0 ? KT : EXT_MAX_LEVELS;
    const size_t total = (size_t)nx * ny * nz;
    const size_t ys = nx, zs = (size_t)nx * ny;
    const size_t base = (size_t)blockIdx.x * EXT_CHUNK;
    __shared__ int s_cnt[EXT_MAX_LEVELS];
    if (threadIdx.x < EXT_MAX_LEVELS) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    float thr[KA];
#pragma unroll
    for (int s = 0; s < KA; s++)  // thr = (float)(peak_thresh * dogmax), sift.c:1169
        thr[s] = s < K ? (float)(peak_thresh * (double)__uint_as_float(L.maxbits[s + 1])) : 0.0f;
    int cnt[KA];
#pragma unroll
    for (int s = 0; s < KA; s++) cnt[s] = 0;
    for (int it = 0; it < EXT_CHUNK / EXT_BLOCK; it += EXT_UNROLL) {
        size_t idx[EXT_UNROLL];
        bool interior[EXT_UNROLL];
        float v[EXT_UNROLL][KA];
#pragma unroll
        for (int u = 0; u < EXT_UNROLL; u++) {
            idx[u] = base + (size_t)(it + u) * EXT_BLOCK + threadIdx.x;
            interior[u] = false;
            if (idx[u] < total) {
                int x, y, z;
                if (total < 0xffffffffull) {  // 32-bit divisions
                    const unsigned i32 = (unsigned)idx[u];
                    const unsigned r = i32 / (unsigned)nx;
                    x = (int)(i32 - r * (unsigned)nx);
                    z = (int)(r / (unsigned)ny);
                    y = (int)(r - (unsigned)z * (unsigned)ny);
                } else {
                    x = (int)(idx[u] % nx);
                    const size_t r = idx[u] / nx;
                    y = (int)(r % ny);
                    z = (int)(r / ny);
                }
                interior[u] = x >= 1 && x <= nx - 2 && y >= 1 && y <= ny - 2 && z >= 1 && z <= nz - 2;
            }
#pragma unroll
            for (int s = 0; s < KA; s++)
                v[u][s] = (interior[u] && s < K) ? __ldg(L.dog[s + 1] + idx[u]) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < EXT_UNROLL; u++) {
            if (base + (size_t)(it + u) * EXT_BLOCK >= total) break;  // block-uniform
            // The eight neighbours of every level that passes the threshold are fetched with
            // PREDICATED loads issued back to back (one latency for all levels of the voxel,
            // not one dependent branch per level); compares and ballots follow.
            float nb[KA][8];
            bool need[KA];
#pragma unroll
            for (int s = 0; s < KA; s++) {
                need[s] = s < K && interior[u] && (v[u][s] > thr[s] || v[u][s] < -thr[s]);
                const float *cur = L.dog[s < K ? s + 1 : 1] + idx[u];
                nb[s][0] = need[s] ? __ldg(L.dog[s < K ? s : 0] + idx[u]) : 0.0f;
                nb[s][1] = need[s] ? __ldg(L.dog[s < K ? s + 2 : 2] + idx[u]) : 0.0f;
                nb[s][2] = need[s] ? __ldg(cur + 1) : 0.0f;
                nb[s][3] = need[s] ? __ldg(cur - 1) : 0.0f;
                nb[s][4] = need[s] ? __ldg(cur + ys) : 0.0f;
                nb[s][5] = need[s] ? __ldg(cur - ys) : 0.0f;
                nb[s][6] = need[s] ? __ldg(cur - zs) : 0.0f;
                nb[s][7] = need[s] ? __ldg(cur + zs) : 0.0f;
            }
#pragma unroll
            for (int s = 0; s < KA; s++) {
                if (s >= K) break;
                const float c = v[u][s];
                bool mx = need[s], mn = need[s];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    mx = mx && c > nb[s][j];
                    mn = mn && c < nb[s][j];
                }
                const unsigned m = __ballot_sync(0xffffffffu, mx || mn);
                if ((threadIdx.x & 31) == 0) {
                    mask[(size_t)s * words_per_level + (idx[u] >> 5)] = m;
                    cnt[s] += __popc(m);
                }
            }
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int s = 0; s < KA; s++)
            if (s < K && cnt[s]) atomicAdd(&s_cnt[s], cnt[s]);
    }
    __syncthreads();
    if (threadIdx.x < K) blockcnt[(size_t)threadIdx.x * nblocks + blockIdx.x] = s_cnt[threadIdx.x];
}

// Pass A for row lengths that are a multiple of 4 (K == 3): a thread owns FOUR consecutive
// voxels of one row -- one aligned 16-byte load per level and neighbour direction instead of
// four 4-byte ones, the x neighbours come out of the same registers -- and the strict
// 8-neighbour tests are one max and one min over the neighbours (FMNMX3) and two compares.
// Eight lanes assemble a 32-voxel mask word with three shuffles.  Same masks, same counts.
__global__ void __launch_bounds__(EXT_BLOCK, 3)
    k_extrema_mark4(const ExtLevels L, int nx, int ny, int nz, double peak_thresh,
                    unsigned *__restrict__ mask, int *__restrict__ blockcnt, int nblocks,
                    size_t words_per_level)
{
    constexpr int K = 3;
    const size_t total = (size_t)nx * ny * nz;
    const size_t ys = nx, zs = (size_t)nx * ny;
    const size_t base = (size_t)blockIdx.x * EXT_CHUNK;
    __shared__ int s_cnt[EXT_MAX_LEVELS];
    if (threadIdx.x < EXT_MAX_LEVELS) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    float thr[K];
#pragma unroll
    for (int s = 0; s < K; s++)  // thr = (float)(peak_thresh * dogmax), sift.c:1169
        thr[s] = (float)(peak_thresh * (double)__uint_as_float(L.maxbits[s + 1]));
    int cnt[K] = {0, 0, 0};
    const int lane = threadIdx.x & 31;
    for (int it = 0; it < EXT_CHUNK / (4 * EXT_BLOCK); it++) {
        const size_t idx = base + 4 * ((size_t)it * EXT_BLOCK + threadIdx.x);  // first of 4 voxels
        if (base + 4 * (size_t)it * EXT_BLOCK >= total) break;                // block-uniform
        bool row_in = false;
        int x = 0;
        if (idx < total) {
            int y, z;
            if (total < 0xffffffffull) {  // 32-bit divisions
                const unsigned i32 = (unsigned)idx;
                const unsigned r = i32 / (unsigned)nx;
                x = (int)(i32 - r * (unsigned)nx);
                z = (int)(r / (unsigned)ny);
                y = (int)(r - (unsigned)z * (unsigned)ny);
            } else {
                x = (int)(idx % nx);
                const size_t r = idx / nx;
                y = (int)(r % ny);
                z = (int)(r / ny);
            }
            row_in = y >= 1 && y <= ny - 2 && z >= 1 && z <= nz - 2;
        }
        float4 v[K];
#pragma unroll
        for (int s = 0; s < K; s++)
            v[s] = row_in ? __ldg(reinterpret_cast<const float4 *>(L.dog[s + 1] + idx))
                          : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int s = 0; s < K; s++) {
            const float c[4] = {v[s].x, v[s].y, v[s].z, v[s].w};
            bool pass[4];
            bool any = false;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                pass[k] = row_in && x + k >= 1 && x + k <= nx - 2 && (c[k] > thr[s] || c[k] < -thr[s]);
                any = any || pass[k];
            }
            unsigned nib = 0;
            if (any) {
                const float *cur = L.dog[s + 1] + idx;
                const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 p = __ldg(reinterpret_cast<const float4 *>(L.dog[s] + idx));
                const float4 q = __ldg(reinterpret_cast<const float4 *>(L.dog[s + 2] + idx));
                const float4 yp = __ldg(reinterpret_cast<const float4 *>(cur + ys));
                const float4 ym = __ldg(reinterpret_cast<const float4 *>(cur - ys));
                const float4 zp = __ldg(reinterpret_cast<const float4 *>(cur + zs));
                const float4 zm = __ldg(reinterpret_cast<const float4 *>(cur - zs));
                const float xl = x > 0 ? __ldg(cur - 1) : 0.0f;          // only used by voxel 0
                const float xr = x + 4 < nx ? __ldg(cur + 4) : 0.0f;     // only used by voxel 3
                (void)z4;
                const float P[4] = {p.x, p.y, p.z, p.w}, Q[4] = {q.x, q.y, q.z, q.w};
                const float YP[4] = {yp.x, yp.y, yp.z, yp.w}, YM[4] = {ym.x, ym.y, ym.z, ym.w};
                const float ZP[4] = {zp.x, zp.y, zp.z, zp.w}, ZM[4] = {zm.x, zm.y, zm.z, zm.w};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const float l = k == 0 ? xl : c[k - 1], r = k == 3 ? xr : c[k + 1];
                    // strict extremum over the 6 + 1 + 1 neighbours (sift.c:1184-1190): c > all
                    // <=> c > max, c < all <=> c < min (no NaNs in a DoG of finite data)
                    const float mx = fmaxf(fmaxf(fmaxf(P[k], Q[k]), fmaxf(l, r)),
                                           fmaxf(fmaxf(YP[k], YM[k]), fmaxf(ZP[k], ZM[k])));
                    const float mn = fminf(fminf(fminf(P[k], Q[k]), fminf(l, r)),
                                           fminf(fminf(YP[k], YM[k]), fminf(ZP[k], ZM[k])));
                    if (pass[k] && (c[k] > mx || c[k] < mn)) nib |= 1u << k;
                }
            }
            // eight lanes (32 voxels) -> one mask word
            unsigned w = nib << (4 * (lane & 7));
            w |= __shfl_xor_sync(0xffffffffu, w, 1);
            w |= __shfl_xor_sync(0xffffffffu, w, 2);
            w |= __shfl_xor_sync(0xffffffffu, w, 4);
            if ((lane & 7) == 0) {
                mask[(size_t)s * words_per_level + (idx >> 5)] = w;
                cnt[s] += __popc(w);
            }
        }
    }
    if ((lane & 7) == 0) {
#pragma unroll
        for (int s = 0; s < K; s++)
            if (cnt[s]) atomicAdd(&s_cnt[s], cnt[s]);
    }
    __syncthreads();
    if (threadIdx.x < K) blockcnt[(size_t)threadIdx.x * nblocks + blockIdx.x] = s_cnt[threadIdx.x];
}

// Pass B: exclusive scan of the per-block counts, level after level, continuing
// the running candidate total kept in counter[0].  One block.
__global__ void __launch_bounds__(1024) k_extrema_scan(int *__restrict__ blockcnt, int nblocks,
                                                      int K, int *__restrict__ counter)
{
    __shared__ int s_warp[32];
    __shared__ int s_run;
    if (threadIdx.x == 0) s_run = counter[0];
    __syncthreads();
    const size_t total = (size_t)K * nblocks;
    for (size_t base = 0; base < total; base += 1024) {
        const size_t i = base + threadIdx.x;
        const int v = i < total ? blockcnt[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += t;
        }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = s_warp[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, w, o);
                if (threadIdx.x >= o) w += t;
            }
            s_warp[threadIdx.x] = w;
        }
        __syncthreads();
        const int warp_off = (threadIdx.x >> 5) ? s_warp[(threadIdx.x >> 5) - 1] : 0;
        const int run = s_run;
        if (i < total) blockcnt[i] = run + warp_off + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_run = run + warp_off + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) counter[0] = s_run;
}

// Pass C: ordered scatter of the marked voxels.  One thread per mask word; a block covers the
// EXT_CHUNK voxels (EXT_BLOCK words) of one pass-A block, whose offset pass B left in blockoff.
__global__ void __launch_bounds__(EXT_BLOCK)
    k_extrema_emit(const unsigned *__restrict__ mask, const int *__restrict__ blockoff, int nblocks,
                   size_t words_per_level, int K, int o, int nx, int ny, int nz, int zbase,
                   Candidate *__restrict__ cand, int cap)
{
    const size_t total = (size_t)nx * ny * nz;
    const size_t word = (size_t)blockIdx.x * EXT_BLOCK + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ int s_w[EXT_BLOCK / 32];
    for (int s = 0; s < K; s++) {
        unsigned m = (word * 32 < total) ? mask[(size_t)s * words_per_level + word] : 0u;
        const int v = __popc(m);
        int incl = v;
#pragma unroll
        for (int k = 1; k < 32; k <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, k);
            if (lane >= k) incl += t;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        int pos = blockoff[(size_t)s * nblocks + blockIdx.x] + incl - v;
        for (int w = 0; w < warp; w++) pos += s_w[w];
        while (m) {
            const int bit = __ffs(m) - 1;
            m &= m - 1;
            if (pos < cap) {
                const size_t idx = word * 32 + bit;
                Candidate c;
                c.o = (short)o;
                c.s = (short)s;
                c.x = (int)(idx % nx);
                const size_t r = idx / nx;
                c.y = (int)(r % ny);
                c.z = (int)(r / ny) + zbase;
                cand[pos] = c;
            }
            pos++;
        }
        __syncthreads();
    }
}

// leading scalars before a common 16-byte boundary; n (all scalar) if the pointers disagree
inline size_t head_of(size_t n, const void *a, const void *b = nullptr, const void *c = nullptr)
{
    const size_t ma = (size_t)((uintptr_t)a & 15);
    if ((b && ((uintptr_t)b & 15) != ma) || (c && ((uintptr_t)c & 15) != ma) || (ma & 3)) return n;
    return std::min(n, ((16 - ma) & 15) / 4);
}

// uf == 2^-k (k >= 0) and every coordinate i - d*uf, i < n, exact in f32
inline int is_dyadic(float uf, int n)
{
    int ex;
    if (!(uf > 0.0f) || frexpf(uf, &ex) != 0.5f || ex > 1) return 0;  // uf = 2^(ex-1)
    const int k = 1 - ex;
    return k <= 20 && (long long)n < (1ll << (24 - k));
}

inline int grid_for(const s3d_engine *e, size_t work_items, int block, int per_sm)
{
    const size_t want = (work_items + block - 1) / block;
    const size_t cap = (size_t)e->num_sms * per_sm;
    return (int)(want < cap ? (want ? want : 1) : cap);
}

}  // namespace

int s3d_k_max_abs(s3d_engine *e, const float *x, size_t n, unsigned *d_bits)
{
    S3D_CUDA(e, cudaMemsetAsync(d_bits, 0, sizeof(unsigned), e->stream));
    k_max_abs<<<grid_for(e, n / 4 + 1, 256, 8), 256, 0, e->stream>>>(x, n, head_of(n, x), d_bits);
    S3D_LAUNCH_CHECK(e);
    return 0;
}

int s3d_k_scale(s3d_engine *e, const float *src, float *dst, size_t n, const unsigned *d_bits)
{
    k_scale<<<grid_for(e, n / 4 + 1, 256, 8), 256, 0, e->stream>>>(src, dst, n,
                                                                   head_of(n, src, dst), d_bits);
    S3D_LAUNCH_CHECK(e);
    return 0;
}

int s3d_ensure_scratch(s3d_engine *e, size_t elems)
{
    if (elems <= e->scratch_cap) return 0;
    for (int i = 0; i < 2; i++) {
        if (e->scratch[i]) cudaFree(e->scratch[i]);
        e->scratch[i] = nullptr;
    }
    e->scratch_cap = 0;
    for (int i = 0; i < 2; i++) S3D_CUDA(e, cudaMalloc(&e->scratch[i], elems * sizeof(float)));
    e->scratch_cap = elems;
    return 0;
}

int s3d_blur_fused(s3d_engine *e, const float *src, float *dst, int nx, int ny, int nz,
                   const TapSet &taps, const float uf[3]);  // blur_fused.cu
bool s3d_blur_fused_eligible(int nx, int ny, int nz, int nc, const TapSet &taps,
                             const float uf[3]);
int s3d_conv_dyadic_order(const TapSet &taps, float uf, int n);  // blur_dyadic.cu
int s3d_conv_dyadic_axis(s3d_engine *e, int axis, int order, const float *src, float *dst, int nx,
                         int ny, int nz, const TapSet &taps);
int s3d_conv_dyadic_zrange(s3d_engine *e, int order, const float *src, float *dst, int nx, int ny,
                           int nbuf, int zb, int ze, int gbase, int nglob, const TapSet &taps);

int s3d_k_blur(s3d_engine *e, const float *src, float *dst, int nx, int ny, int nz, int nc,
               const TapSet &taps, const float uf[3])
{
    if (e->blur_mode == 0 && s3d_blur_fused_eligible(nx, ny, nz, nc, taps, uf))
        return s3d_blur_fused(e, src, dst, nx, ny, nz, taps, uf);
    const size_t total = (size_t)nx * ny * nz * nc;
    if (s3d_ensure_scratch(e, total)) return -1;
    const int grid = grid_for(e, total, 256, 16);
    // dyadic tap spacings (octaves 0-2 of a pyramid with power-of-two units, the 12-channel dense
    // window blur): register-blocked per-axis kernels (blur_dyadic.cu).  Interleaved channels are
    // a reinterpretation of the same layout: along x the array is [ny*nz rows][nx][nc] -- a "y"
    // pass over lines of stride nc -- and along y / z a row is nx*nc contiguous floats.
    const int dims[3] = {nx, ny, nz};
    int ord[3] = {-1, -1, -1};
    if (e->blur_mode == 0 && (size_t)ny * nz < 0x7fffffffull && (size_t)nx * nc < 0x7fffffffull)
        for (int a = 0; a < 3; a++) ord[a] = s3d_conv_dyadic_order(taps, uf[a], dims[a]);
    int rc = 1;
    if (ord[0] >= 0)
        rc = nc == 1 ? s3d_conv_dyadic_axis(e, 0, ord[0], src, e->scratch[0], nx, ny, nz, taps)
                     : s3d_conv_dyadic_axis(e, 1, ord[0], src, e->scratch[0], nc, nx, ny * nz, taps);
    if (rc < 0) return -1;
    if (rc) {
        k_conv_axis<0><<<grid, 256, 0, e->stream>>>(src, e->scratch[0], nx, ny, nz, nc, taps, uf[0],
                                                   0, nz, 0, nz, is_dyadic(uf[0], nx));
        S3D_LAUNCH_CHECK(e);
    }
    rc = 1;
    if (ord[1] >= 0)
        rc = s3d_conv_dyadic_axis(e, 1, ord[1], e->scratch[0], e->scratch[1], nx * nc, ny, nz, taps);
    if (rc < 0) return -1;
    if (rc) {
        k_conv_axis<1><<<grid, 256, 0, e->stream>>>(e->scratch[0], e->scratch[1], nx, ny, nz, nc,
                                                   taps, uf[1], 0, nz, 0, nz, is_dyadic(uf[1], ny));
        S3D_LAUNCH_CHECK(e);
    }
    rc = 1;
    if (ord[2] >= 0)
        rc = s3d_conv_dyadic_axis(e, 2, ord[2], e->scratch[1], dst, nx * nc, ny, nz, taps);
    if (rc < 0) return -1;
    if (rc) {
        k_conv_axis<2><<<grid, 256, 0, e->stream>>>(e->scratch[1], dst, nx, ny, nz, nc, taps, uf[2],
                                                   0, nz, 0, nz, is_dyadic(uf[2], nz));
        S3D_LAUNCH_CHECK(e);
    }
    return 0;
}

int s3d_blur_z_reach(const TapSet &taps, float ufz)
{  // samples reach c = i +- hw*uf, read as the pair (floor(c), floor(c) + 1)
    return (int)ceilf((float)(taps.width / 2) * ufz) + 1;
}

int s3d_blur_fused_zrange(s3d_engine *e, const float *src, float *dst, int nx, int ny, int nz,
                          const TapSet &taps, int zb, int ze, int gz0, int nz_glob);  // blur_fused.cu

int s3d_k_blur_zrange(s3d_engine *e, const float *src, float *dst, int nx, int ny, int nz,
                      const TapSet &taps, const float uf[3], int zb, int ze, int gz0, int nz_glob)
{
    if (ze <= zb) return 0;
    if (e->blur_mode == 0 && s3d_blur_fused_eligible(nx, ny, nz, 1, taps, uf))
        return s3d_blur_fused_zrange(e, src, dst, nx, ny, nz, taps, zb, ze, gz0, nz_glob);
    // x and y passes on the planes the z pass will read, then the z pass on [zb, ze).  The
    // scratch volumes hold planes [p0, p1) only; the z pass computes its sample coordinates
    // with GLOBAL plane indices (buffer plane 0 = global plane gz0) so that every f32
    // rounding -- incl. the mirror at the true end of the volume -- is the whole-volume one.
    const int h = s3d_blur_z_reach(taps, uf[2]);
    const int p0 = std::max(0, zb - h), p1 = std::min(nz, ze + h);
    const size_t plane = (size_t)nx * ny;
    const size_t sub = plane * (size_t)(p1 - p0);
    if (s3d_ensure_scratch(e, sub)) return -1;
    const int grid = grid_for(e, sub, 256, 16);
    // octaves 1 and 2 of a dyadic pyramid: the register-blocked kernels (blur_dyadic.cu); x and y
    // are plain passes over the planes [p0, p1), z takes the plane-range variant
    int ord[3] = {-1, -1, -1};
    if (e->blur_mode == 0) {
        ord[0] = s3d_conv_dyadic_order(taps, uf[0], nx);
        ord[1] = s3d_conv_dyadic_order(taps, uf[1], ny);
        ord[2] = s3d_conv_dyadic_order(taps, uf[2], nz_glob);
    }
    int rc = 1;
    if (ord[0] >= 1)
        rc = s3d_conv_dyadic_axis(e, 0, ord[0], src + plane * p0, e->scratch[0], nx, ny, p1 - p0, taps);
    if (rc < 0) return -1;
    if (rc) {
        k_conv_axis<0><<<grid, 256, 0, e->stream>>>(src + plane * p0, e->scratch[0], nx, ny, p1 - p0,
                                                   1, taps, uf[0], 0, 0, 0, 0, is_dyadic(uf[0], nx));
        S3D_LAUNCH_CHECK(e);
    }
    rc = 1;
    if (ord[1] >= 1)
        rc = s3d_conv_dyadic_axis(e, 1, ord[1], e->scratch[0], e->scratch[1], nx, ny, p1 - p0, taps);
    if (rc < 0) return -1;
    if (rc) {
        k_conv_axis<1><<<grid, 256, 0, e->stream>>>(e->scratch[0], e->scratch[1], nx, ny, p1 - p0, 1,
                                                   taps, uf[1], 0, 0, 0, 0, is_dyadic(uf[1], ny));
        S3D_LAUNCH_CHECK(e);
    }
    rc = 1;
    if (ord[2] >= 1)  // scratch[1] holds planes [p0, p1) = global [gz0 + p0, gz0 + p1)
        rc = s3d_conv_dyadic_zrange(e, ord[2], e->scratch[1], dst + plane * p0, nx, ny, p1 - p0,
                                    zb - p0, ze - p0, gz0 + p0, nz_glob, taps);
    if (rc < 0) return -1;
    if (rc) {
        k_conv_axis<2><<<grid_for(e, plane * (size_t)(ze - zb), 256, 16), 256, 0, e->stream>>>(
            e->scratch[1], dst + plane * zb, nx, ny, ze - zb, 1, taps, uf[2], zb - p0, p1 - p0,
            gz0 + p0, nz_glob, is_dyadic(uf[2], nz_glob));
        S3D_LAUNCH_CHECK(e);
    }
    return 0;
}

int s3d_k_decimate(s3d_engine *e, const float *src, int sx, int sy, int sz, float *dst, int dx,
                   int dy, int dz)
{
    (void)sz;
    const size_t total = (size_t)dx * dy * dz;
    k_decimate<<<grid_for(e, total, 256, 16), 256, 0, e->stream>>>(src, sx, sy, dst, dx, dy, dz);
    S3D_LAUNCH_CHECK(e);
    return 0;
}

int s3d_k_dog(s3d_engine *e, const float *a, const float *b, float *d, size_t n,
              unsigned *d_maxbits)
{
    k_dog<<<grid_for(e, n / 4 + 1, 256, 8), 256, 0, e->stream>>>(a, b, d, n, head_of(n, a, b, d),
                                                                 d_maxbits);
    S3D_LAUNCH_CHECK(e);
    return 0;
}

// build_dog for one octave (levels consecutive in e->g / e->dog); returns 1 if the fused kernel
// does not apply (the caller then runs s3d_k_dog per level)
int s3d_k_dog_octave(s3d_engine *e, int o)
{
    const int nl = e->nlev_g;
    if (nl < 2 || nl > DOG_MAX_LEVELS || e->nlev_d != nl - 1) return 1;
    const size_t n = e->g[(size_t)o * nl].n();
    if (n == 0 || (n & 3)) return 1;
    DogLevels L;
    L.nl = nl;
    for (int s = 0; s < nl; s++) {
        L.g[s] = e->g[(size_t)o * nl + s].d;
        if (((uintptr_t)L.g[s] & 15) || e->g[(size_t)o * nl + s].n() != n) return 1;
    }
    for (int s = 0; s < nl - 1; s++) {
        L.d[s] = e->dog[(size_t)o * e->nlev_d + s].d;
        if (((uintptr_t)L.d[s] & 15) || e->dog[(size_t)o * e->nlev_d + s].n() != n) return 1;
    }
    L.maxbits = e->d_scalars + 1 + (size_t)o * e->nlev_d;
    const int grid = grid_for(e, n / 4, 256, 8);
    if (nl == 6)
        k_dog_octave<6><<<grid, 256, 0, e->stream>>>(L, n / 4);
    else
        k_dog_octave<0><<<grid, 256, 0, e->stream>>>(L, n / 4);
    S3D_LAUNCH_CHECK(e);
    return 0;
}

int s3d_k_extrema_octave(s3d_engine *e, int o, float, double peak_thresh)
{
    return s3d_k_extrema_range(e, o, peak_thresh, 0, e->dog[(size_t)o * e->nlev_d].g.nz, 0);
}

int s3d_k_extrema_range(s3d_engine *e, int o, double peak_thresh, int zl0, int nzs, int zbase)
{
    const LevelDev &l0 = e->dog[(size_t)o * e->nlev_d];
    const int nx = l0.g.nx, ny = l0.g.ny, nz = nzs;
    const size_t zskip = (size_t)nx * ny * zl0;
    const size_t total = (size_t)nx * ny * nz;
    const int nblocks = (int)((total + EXT_CHUNK - 1) / EXT_CHUNK);
    const size_t words = (size_t)nblocks * (EXT_CHUNK / 32);
    const int K = e->K;
    if (K > EXT_MAX_LEVELS) return s3d_fail(e, "num_kp_levels too large", cudaSuccess, __FILE__, __LINE__);
    if (words * K > e->mask_cap) {
        if (e->d_mask) cudaFree(e->d_mask);
        e->d_mask = nullptr;
        e->mask_cap = 0;
        S3D_CUDA(e, cudaMalloc(&e->d_mask, words * K * sizeof(unsigned)));
        e->mask_cap = words * K;
    }
    if ((size_t)nblocks * K > e->blockcnt_cap) {
        if (e->d_blockcnt) cudaFree(e->d_blockcnt);
        e->d_blockcnt = nullptr;
        e->blockcnt_cap = 0;
        S3D_CUDA(e, cudaMalloc(&e->d_blockcnt, (size_t)nblocks * K * sizeof(int)));
        e->blockcnt_cap = (size_t)nblocks * K;
    }
    ExtLevels L;
    for (int s = -1; s <= K; s++) L.dog[s + 1] = e->dog[(size_t)o * e->nlev_d + (s + 1)].d + zskip;
    L.maxbits = e->d_scalars + 1 + (size_t)o * e->nlev_d;
    L.K = K;
    bool vec4 = K == 3 && (nx & 3) == 0 && e->blur_mode == 0;
    for (int s = 0; s <= K + 1 && vec4; s++) vec4 = ((uintptr_t)L.dog[s] & 15) == 0;
    if (vec4)
        k_extrema_mark4<<<nblocks, EXT_BLOCK, 0, e->stream>>>(L, nx, ny, nz, peak_thresh, e->d_mask,
                                                             e->d_blockcnt, nblocks, words);
    else if (K == 3)
        k_extrema_mark<3><<<nblocks, EXT_BLOCK, 0, e->stream>>>(L, nx, ny, nz, peak_thresh, e->d_mask,
                                                               e->d_blockcnt, nblocks, words);
    else
        k_extrema_mark<0><<<nblocks, EXT_BLOCK, 0, e->stream>>>(L, nx, ny, nz, peak_thresh, e->d_mask,
                                                               e->d_blockcnt, nblocks, words);
    S3D_LAUNCH_CHECK(e);
    k_extrema_scan<<<1, 1024, 0, e->stream>>>(e->d_blockcnt, nblocks, K, e->d_counter);
    S3D_LAUNCH_CHECK(e);
    k_extrema_emit<<<nblocks, EXT_BLOCK, 0, e->stream>>>(e->d_mask, e->d_blockcnt, nblocks, words,
                                                        K, o, nx, ny, nz, zbase, e->d_cand,
                                                        e->cand_cap);
    S3D_LAUNCH_CHECK(e);
    return 0;
}
