// blur_dyadic.cu -- per-axis Gaussian FIR for dyadic tap spacings 2^-O, O = 0, 1, 2 (pyramid
// octaves 0-2 of a volume with power-of-two units), register-blocked.
//
// Replaces convolve_sep_gen (imutil.c:2274-2393) for those octaves; same arithmetic contract as
// k_conv_axis (pyramid.cu): acc = acc (+) tap (*) ((1-frac) (*) lo (+) frac (*) hi), one IEEE
// rounding per operation, taps visited d = -hw..hw.
//
// With spacing 2^-O an interior output i samples c = i - d * 2^-O, i.e. the lerped value
//     L_r[j] = (1 - r/P) * s[j] + (r/P) * s[j + 1],   P = 2^O,  j = i + floor(-d / P),  r = (-d) mod P
// which does not depend on i: neighbouring outputs share it.  A thread owns RUN consecutive
// outputs of one line and walks the samples from HIGH to LOW coordinate (j descending, r
// descending within j): every L_r[j] is formed once (3 operations) and scattered into the
// accumulators of the outputs that use it.  Output i receives its taps in the order of
// decreasing c = increasing d -- the reference's order -- so the result is bit-identical, at
// 2 FP32 operations per tap instead of 5 plus two loads.  Everything is unrolled: tap, phase and
// accumulator indices are compile-time constants (taps are read from the constant bank).
// Outputs outside the reference's interior range [uhw, n-2-uhw] are recomputed with the literal
// mirror path; whole runs past the volume end are skipped.
#include "common.cuh"

#include <cmath>

namespace {

// The per-line arithmetic is __host__ __device__ so that tools/dyadic_host_check.cu can run the
// very same code on the CPU against the oracle (x86-64 baseline has no FMA: a*b+c is two roundings).
#ifdef __CUDA_ARCH__
#define RMUL(a, b) __fmul_rn((a), (b))
#define RADD(a, b) __fadd_rn((a), (b))
#define RSUB(a, b) __fsub_rn((a), (b))
#define LDG(p) __ldg(p)
#define F2I_RZ(c) __float2int_rz(c)
#else
static inline float h_mul(float a, float b) { volatile float r = a * b; return r; }
static inline float h_add(float a, float b) { volatile float r = a + b; return r; }
static inline float h_sub(float a, float b) { volatile float r = a - b; return r; }
#define RMUL(a, b) h_mul((a), (b))
#define RADD(a, b) h_add((a), (b))
#define RSUB(a, b) h_sub((a), (b))
#define LDG(p) (*(p))
#define F2I_RZ(c) ((int)(c))
#endif

__host__ __device__ __forceinline__ int s3d_clampi(int v, int hi) { return v < 0 ? 0 : (v > hi ? hi : v); }

__host__ __device__ __forceinline__ float lit_samp(float acc, float tap, float c, const float *line,
                                          size_t st, int dim_end)
{
    int lo = F2I_RZ(c);
    const float frac = RSUB(c, (float)lo);
    int hi = lo + 1;
    lo = s3d_clampi(lo, dim_end);
    hi = s3d_clampi(hi, dim_end);
    const float v = RADD(RMUL(RSUB(1.0f, frac), LDG(line + (size_t)lo * st)),
                              RMUL(frac, LDG(line + (size_t)hi * st)));
    return RADD(acc, RMUL(tap, v));
}

// boundary pass of convolve_sep_gen (imutil.c:2365-2387)
__host__ __device__ __noinline__ float boundary_point(const float *line, size_t st, int n, int i,
                                             const TapSet &taps, float uf)
{
    const int hw = taps.width / 2;
    const int dim_end = n - 1;
    float acc = 0.0f;
    for (int d = -hw; d <= hw; d++) {
        const float step = RMUL((float)d, uf);
        float c = RSUB((float)i, step);
        if (F2I_RZ(c) < 0)
            c = -c;
        else if (F2I_RZ(c) >= dim_end)
            c = RSUB(RSUB(RMUL(2.0f, (float)dim_end), c), 0.1f);
        acc = lit_samp(acc, taps.t[d + hw], c, line, st, dim_end);
    }
    return acc;
}

// ld(jj) returns sample i0 + jj of the line (any jj in [-UHW, RUN + OFFMAX])
template <int O, int HW, int RUN, class Loader>
__host__ __device__ __forceinline__ void conv_run_ld(const Loader &ld, const TapSet &taps,
                                                     float (&acc)[RUN])
{
    constexpr int P = 1 << O;
    constexpr int UHW = (HW + P - 1) >> O;  // ceil(HW / P): reach towards lower coordinates
    constexpr int OFFMAX = HW >> O;         // floor(HW / P): reach towards higher coordinates
#pragma unroll
    for (int k = 0; k < RUN; k++) acc[k] = 0.0f;
    float hi = ld(RUN + OFFMAX);
#pragma unroll
    for (int jj = RUN - 1 + OFFMAX; jj >= -UHW; jj--) {  // sample j = i0 + jj
        const float lo = ld(jj);
#pragma unroll
        for (int r = P - 1; r >= 0; r--) {
            // used by output k iff d = -((jj - k) * P + r) lies in [-HW, HW]
            bool used = false;
#pragma unroll
            for (int k = 0; k < RUN; k++) {
                const int d = -((jj - k) * P + r);
                used = used || (d >= -HW && d <= HW);
            }
            if (!used) continue;
            const float f = (float)r / (float)P;  // exact
            const float L = RADD(RMUL(1.0f - f, lo), RMUL(f, hi));
#pragma unroll
            for (int k = 0; k < RUN; k++) {
                const int d = -((jj - k) * P + r);
                if (d >= -HW && d <= HW) acc[k] = RADD(acc[k], RMUL(taps.t[d + HW], L));
            }
        }
        hi = lo;
    }
}

template <int O, int HW, int RUN>
__host__ __device__ __forceinline__ void conv_run(const float *__restrict__ line, size_t st, int n,
                                                  int i0, const TapSet &taps, float (&acc)[RUN])
{
    const int last = n - 1;
    conv_run_ld<O, HW, RUN>(
        [&](int jj) { return LDG(line + (size_t)s3d_clampi(i0 + jj, last) * st); }, taps, acc);
}

// ---- x axis through shared memory -------------------------------------------------------------
// A block stages XT_ROWS rows x (XT_OUT outputs + filter reach) samples with coalesced loads,
// every thread then owns 8 consecutive outputs of one staged row (32 threads per row), and the
// results go back through the same buffer for coalesced stores.  One pad word per 8 samples
// makes the lanes' strided accesses (8 words apart) conflict-free: word c lives at c + (c >> 3),
// lane l reads 9 * l + const.  The three phases are __host__ __device__ (tools/dyadic_host_check.cu
// runs them tid by tid); the kernel separates them with __syncthreads().
#define XT_OUT 256
#define XT_ROWS 8
#define XT_RUN 8

template <int O, int HW>
struct XTile {
    static constexpr int P = 1 << O;
    static constexpr int UHW = (HW + P - 1) >> O;
    static constexpr int OFFMAX = HW >> O;
    static constexpr int W = XT_OUT + UHW + OFFMAX + 1;  // staged samples per row
    static constexpr int WP = W + (W >> 3) + 1;           // with pad words
};

__host__ __device__ __forceinline__ int xt_phys(int c) { return c + (c >> 3); }

template <int O, int HW>
__host__ __device__ __forceinline__ void xtile_load(int tid, float *s, const float *src, int nx,
                                                    size_t nrows, int x_base, size_t row0)
{
    using T = XTile<O, HW>;
    for (int e = tid; e < XT_ROWS * T::W; e += XT_OUT) {
        const int r = e / T::W, c = e - r * T::W;
        const size_t row = row0 + r;
        if (row < nrows)
            s[r * T::WP + xt_phys(c)] = LDG(src + row * (size_t)nx + s3d_clampi(x_base - T::UHW + c, nx - 1));
    }
}

template <int O, int HW>
__host__ __device__ __forceinline__ void xtile_compute(int tid, const float *s, const float *src,
                                                       int nx, size_t nrows, int x_base, size_t row0,
                                                       const TapSet &taps, float (&acc)[XT_RUN])
{
    using T = XTile<O, HW>;
    const int r = tid >> 5, l = tid & 31;
    const size_t row = row0 + r;
    const int i0 = x_base + XT_RUN * l;
    if (row >= nrows || i0 >= nx) return;
    const float *srow = s + r * T::WP;
    const int c0 = XT_RUN * l + T::UHW;  // staged index of sample i0
    conv_run_ld<O, HW, XT_RUN>([&](int jj) { return srow[xt_phys(c0 + jj)]; }, taps, acc);
}

// Outputs outside the reference's interior range [uhw, nx-2-uhw] (a handful at either end of a
// row) take the literal mirror path.  The 32 lanes that own a staged row share them, one output
// per lane: left to their owners they would serialise 8 outputs x all taps in the first and the
// last lane of EVERY warp while the other 30 lanes wait.
template <int O, int HW>
__host__ __device__ __forceinline__ void xtile_fix_ends(int tid, float *s, const float *src, int nx,
                                                        size_t nrows, int x_base, size_t row0,
                                                        const TapSet &taps)
{
    using T = XTile<O, HW>;
    const int r = tid >> 5, l = tid & 31;
    const size_t row = row0 + r;
    if (row >= nrows) return;
    const int start = T::UHW, end = nx - 1 - (T::UHW + 1);
    const int nl = start < nx ? start : nx;                  // outputs [0, nl)
    const int rb = end + 1 > nl ? end + 1 : nl;               // outputs [rb, nx)
    const int nb = nl + (nx - rb);
    const float *line = src + row * (size_t)nx;
    for (int b = l; b < nb; b += 32) {
        const int i = b < nl ? b : rb + (b - nl);
        if (i >= x_base && i < x_base + XT_OUT)
            s[r * T::WP + xt_phys(i - x_base)] = boundary_point(line, 1, nx, i, taps, 1.0f / (float)T::P);
    }
}

template <int O, int HW>
__host__ __device__ __forceinline__ void xtile_stage(int tid, float *s, const float (&acc)[XT_RUN])
{
    using T = XTile<O, HW>;
    const int r = tid >> 5, l = tid & 31;
#pragma unroll
    for (int k = 0; k < XT_RUN; k++) s[r * T::WP + xt_phys(XT_RUN * l + k)] = acc[k];
}

template <int O, int HW>
__host__ __device__ __forceinline__ void xtile_store(int tid, const float *s, float *dst, int nx,
                                                     size_t nrows, int x_base, size_t row0)
{
    using T = XTile<O, HW>;
    for (int e = tid; e < XT_ROWS * XT_OUT; e += XT_OUT) {
        const int r = e / XT_OUT, c = e - r * XT_OUT;
        const size_t row = row0 + r;
        if (row < nrows && x_base + c < nx) dst[row * (size_t)nx + x_base + c] = s[r * T::WP + xt_phys(c)];
    }
}

template <int O, int HW>
__global__ void __launch_bounds__(XT_OUT) k_conv_dyadic_x(const float *__restrict__ src,
                                                          float *__restrict__ dst, int nx,
                                                          size_t nrows, int ntx,
                                                          const __grid_constant__ TapSet taps)
{
    using T = XTile<O, HW>;
    __shared__ float s[XT_ROWS * T::WP];
    const int x_base = (int)(blockIdx.x % ntx) * XT_OUT;
    const size_t row0 = (size_t)(blockIdx.x / ntx) * XT_ROWS;
    xtile_load<O, HW>(threadIdx.x, s, src, nx, nrows, x_base, row0);
    __syncthreads();
    float acc[XT_RUN];
#pragma unroll
    for (int k = 0; k < XT_RUN; k++) acc[k] = 0.0f;
    xtile_compute<O, HW>(threadIdx.x, s, src, nx, nrows, x_base, row0, taps, acc);
    __syncthreads();
    xtile_stage<O, HW>(threadIdx.x, s, acc);
    __syncthreads();
    if (x_base < T::UHW || x_base + XT_OUT > nx - 1 - (T::UHW + 1)) {  // tile touches a row end
        xtile_fix_ends<O, HW>(threadIdx.x, s, src, nx, nrows, x_base, row0, taps);
        __syncthreads();
    }
    xtile_store<O, HW>(threadIdx.x, s, dst, nx, nrows, x_base, row0);
}

// y / z axes (AXIS 1 / 2): thread = one x, RUN consecutive y / z; lanes are x neighbours, so every
// load and store of a warp is one coalesced segment.  (AXIS 0 also works -- thread = RUN
// consecutive x -- but its strided accesses are slow: the x axis uses k_conv_dyadic_x.)
template <int AXIS, int O, int HW, int RUN>
__global__ void __launch_bounds__(256) k_conv_dyadic(const float *__restrict__ src,
                                                     float *__restrict__ dst, int nx, int ny,
                                                     int nz, const __grid_constant__ TapSet taps)
{
    constexpr int P = 1 << O;
    constexpr int UHW = (HW + P - 1) >> O;
    const int n = AXIS == 0 ? nx : (AXIS == 1 ? ny : nz);
    const int nruns = (n + RUN - 1) / RUN;
    const size_t st = AXIS == 0 ? 1 : (AXIS == 1 ? (size_t)nx : (size_t)nx * ny);
    const size_t nthreads = AXIS == 0 ? (size_t)nruns * ny * nz
                                      : (AXIS == 1 ? (size_t)nx * nruns * nz : (size_t)nx * ny * nruns);
    const int start = UHW, end = n - 1 - (UHW + 1);
    const float uf = 1.0f / (float)P;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < nthreads;
         t += (size_t)gridDim.x * blockDim.x) {
        size_t line_off;
        int i0;
        if (AXIS == 0) {
            i0 = (int)(t % nruns) * RUN;
            line_off = (t / nruns) * (size_t)nx;
        } else if (AXIS == 1) {
            const int x = (int)(t % nx);
            const size_t r = t / nx;
            i0 = (int)(r % nruns) * RUN;
            line_off = (r / nruns) * (size_t)nx * ny + x;
        } else {
            const size_t plane = (size_t)nx * ny;
            i0 = (int)(t / plane) * RUN;
            line_off = t % plane;
        }
        const float *line = src + line_off;
        float acc[RUN];
        conv_run<O, HW, RUN>(line, st, n, i0, taps, acc);
        if (i0 < start || i0 + RUN - 1 > end) {
#pragma unroll
            for (int k = 0; k < RUN; k++) {
                const int i = i0 + k;
                if (i < n && (i < start || i > end)) acc[k] = boundary_point(line, st, n, i, taps, uf);
            }
        }
        float *o = dst + line_off + (size_t)i0 * st;
#pragma unroll
        for (int k = 0; k < RUN; k++)
            if (i0 + k < n) o[(size_t)k * st] = acc[k];
    }
}

// ---- z axis over a plane range of a Z-slab buffer ----------------------------------------------
// The buffer holds planes [gbase, gbase + nbuf) of a line of nglob planes; outputs are the buffer
// planes [zb, ze).  Sample coordinates, the interior range and the mirror are those of the GLOBAL
// line (the f32 roundings of the mirror depend on the magnitude of the index, and a slab must
// reproduce the whole-volume result bit for bit); only the addressing is shifted and clamped to
// the buffer (an interior output never reaches outside it: the caller provides the halo).
__host__ __device__ __forceinline__ float lit_samp_g(float acc, float tap, float c, const float *line,
                                                     size_t st, int dim_end, int gbase, int nbuf)
{
    int lo = F2I_RZ(c);
    const float frac = RSUB(c, (float)lo);
    int hi = lo + 1;
    lo = s3d_clampi(s3d_clampi(lo, dim_end) - gbase, nbuf - 1);
    hi = s3d_clampi(s3d_clampi(hi, dim_end) - gbase, nbuf - 1);
    const float v = RADD(RMUL(RSUB(1.0f, frac), LDG(line + (size_t)lo * st)),
                         RMUL(frac, LDG(line + (size_t)hi * st)));
    return RADD(acc, RMUL(tap, v));
}

__host__ __device__ __noinline__ float boundary_point_g(const float *line, size_t st, int nglob, int i,
                                                        const TapSet &taps, float uf, int gbase,
                                                        int nbuf)
{
    const int hw = taps.width / 2;
    const int dim_end = nglob - 1;
    float acc = 0.0f;
    for (int d = -hw; d <= hw; d++) {
        const float step = RMUL((float)d, uf);
        float c = RSUB((float)i, step);
        if (F2I_RZ(c) < 0)
            c = -c;
        else if (F2I_RZ(c) >= dim_end)
            c = RSUB(RSUB(RMUL(2.0f, (float)dim_end), c), 0.1f);
        acc = lit_samp_g(acc, taps.t[d + hw], c, line, st, dim_end, gbase, nbuf);
    }
    return acc;
}

// one thread's RUN outputs starting at buffer plane b0 of the line at `line` (buffer plane 0)
template <int O, int HW, int RUN>
__host__ __device__ __forceinline__ void zrange_run(const float *line, float *out, size_t plane,
                                                    int nbuf, int b0, int ze, int gbase, int nglob,
                                                    const TapSet &taps)
{
    constexpr int P = 1 << O;
    constexpr int UHW = (HW + P - 1) >> O;
    float acc[RUN];
    conv_run_ld<O, HW, RUN>(
        [&](int jj) { return LDG(line + (size_t)s3d_clampi(b0 + jj, nbuf - 1) * plane); }, taps, acc);
    const int i0 = gbase + b0;  // global index of the first output
    const int start = UHW, end = nglob - 1 - (UHW + 1);
    if (i0 < start || i0 + RUN - 1 > end) {
#pragma unroll
        for (int k = 0; k < RUN; k++) {
            const int i = i0 + k;
            if (b0 + k < ze && (i < start || i > end))
                acc[k] = boundary_point_g(line, plane, nglob, i, taps, 1.0f / (float)P, gbase, nbuf);
        }
    }
#pragma unroll
    for (int k = 0; k < RUN; k++)
        if (b0 + k < ze) out[(size_t)(b0 + k) * plane] = acc[k];
}

template <int O, int HW, int RUN>
__global__ void __launch_bounds__(256) k_conv_dyadic_zr(const float *__restrict__ src,
                                                        float *__restrict__ dst, int nx, int ny,
                                                        int nbuf, int zb, int ze, int gbase, int nglob,
                                                        const __grid_constant__ TapSet taps)
{
    const size_t plane = (size_t)nx * ny;
    const size_t nruns = (size_t)(ze - zb + RUN - 1) / RUN;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < plane * nruns;
         t += (size_t)gridDim.x * blockDim.x) {
        const size_t line_off = t % plane;
        const int b0 = zb + (int)(t / plane) * RUN;
        zrange_run<O, HW, RUN>(src + line_off, dst + line_off, plane, nbuf, b0, ze, gbase, nglob, taps);
    }
}

template <int O, int HW>
int launch_zr(s3d_engine *e, const float *src, float *dst, int nx, int ny, int nbuf, int zb, int ze,
              int gbase, int nglob, const TapSet &taps)
{
    constexpr int RUN = 16;
    const size_t nthreads = (size_t)nx * ny * ((size_t)(ze - zb + RUN - 1) / RUN);
    const size_t want = (nthreads + 255) / 256;
    const size_t cap = (size_t)e->num_sms * 64;
    const int grid = (int)(want < cap ? (want ? want : 1) : cap);
    k_conv_dyadic_zr<O, HW, RUN><<<grid, 256, 0, e->stream>>>(src, dst, nx, ny, nbuf, zb, ze, gbase,
                                                             nglob, taps);
    S3D_LAUNCH_CHECK(e);
    return 0;
}

template <int O>
int launch_zr_hw(s3d_engine *e, const float *src, float *dst, int nx, int ny, int nbuf, int zb, int ze,
                 int gbase, int nglob, const TapSet &taps)
{
    switch (taps.width / 2) {
    case 3: return launch_zr<O, 3>(e, src, dst, nx, ny, nbuf, zb, ze, gbase, nglob, taps);
    case 4: return launch_zr<O, 4>(e, src, dst, nx, ny, nbuf, zb, ze, gbase, nglob, taps);
    case 5: return launch_zr<O, 5>(e, src, dst, nx, ny, nbuf, zb, ze, gbase, nglob, taps);
    case 6: return launch_zr<O, 6>(e, src, dst, nx, ny, nbuf, zb, ze, gbase, nglob, taps);
    case 8: return launch_zr<O, 8>(e, src, dst, nx, ny, nbuf, zb, ze, gbase, nglob, taps);
    default: return 1;
    }
}

template <int AXIS, int O, int HW>
int launch_one(s3d_engine *e, const float *src, float *dst, int nx, int ny, int nz,
               const TapSet &taps)
{
    if (AXIS == 0) {
        const size_t nrows = (size_t)ny * nz;
        const int ntx = (nx + XT_OUT - 1) / XT_OUT;
        const size_t nblocks = (size_t)ntx * ((nrows + XT_ROWS - 1) / XT_ROWS);
        if (nblocks > 0x7fffffffull) return 1;
        k_conv_dyadic_x<O, HW><<<(unsigned)nblocks, XT_OUT, 0, e->stream>>>(src, dst, nx, nrows, ntx, taps);
        S3D_LAUNCH_CHECK(e);
        return 0;
    }
    constexpr int RUN = 16;
    const int n = AXIS == 1 ? ny : nz;
    const size_t nruns = (size_t)(n + RUN - 1) / RUN;
    const size_t nthreads = (size_t)nx * ny * nz / (size_t)n * nruns;
    const size_t want = (nthreads + 255) / 256;
    const size_t cap = (size_t)e->num_sms * 64;
    const int grid = (int)(want < cap ? (want ? want : 1) : cap);
    k_conv_dyadic<(AXIS == 0 ? 1 : AXIS), O, HW, RUN><<<grid, 256, 0, e->stream>>>(src, dst, nx, ny, nz, taps);
    S3D_LAUNCH_CHECK(e);
    return 0;
}

template <int AXIS, int O>
int launch_hw(s3d_engine *e, const float *src, float *dst, int nx, int ny, int nz,
              const TapSet &taps)
{
    switch (taps.width / 2) {
    case 2: return launch_one<AXIS, O, 2>(e, src, dst, nx, ny, nz, taps);
    case 3: return launch_one<AXIS, O, 3>(e, src, dst, nx, ny, nz, taps);
    case 4: return launch_one<AXIS, O, 4>(e, src, dst, nx, ny, nz, taps);
    case 5: return launch_one<AXIS, O, 5>(e, src, dst, nx, ny, nz, taps);
    case 6: return launch_one<AXIS, O, 6>(e, src, dst, nx, ny, nz, taps);
    case 8: return launch_one<AXIS, O, 8>(e, src, dst, nx, ny, nz, taps);
    case 7: if (O == 0) return launch_one<AXIS, 0, 7>(e, src, dst, nx, ny, nz, taps); return 1;
    case 9: if (O == 0) return launch_one<AXIS, 0, 9>(e, src, dst, nx, ny, nz, taps); return 1;
    case 10: if (O == 0) return launch_one<AXIS, 0, 10>(e, src, dst, nx, ny, nz, taps); return 1;
    default: return 1;
    }
}

}  // namespace

// O with uf = 2^-O, O in {0, 1, 2}, if every coordinate is exact in f32 and this file
// instantiates the half width; -1 otherwise.  (O = 0 serves the volumes the fused kernel of
// blur_fused.cu does not take, e.g. row lengths that are not a multiple of 4.)
int s3d_conv_dyadic_order(const TapSet &taps, float uf, int n)
{
    const int hw = taps.width / 2;
    const bool pyr = hw == 2 || hw == 3 || hw == 4 || hw == 5 || hw == 6 || hw == 8;
    if ((long long)n >= (1ll << 20)) return -1;
    if (uf == 1.0f && (pyr || hw == 7 || hw == 9 || hw == 10)) return 0;  // + dense window widths
    if (!pyr) return -1;
    if (uf == 0.5f) return 1;
    if (uf == 0.25f) return 2;
    return -1;
}

// one axis of the separable filter; returns 1 if (axis, order, width) is not instantiated,
// -1 on a launch error
int s3d_conv_dyadic_axis(s3d_engine *e, int axis, int order, const float *src, float *dst, int nx,
                         int ny, int nz, const TapSet &taps)
{
#define S3D_DY_ORDER(O)                                                        \
    if (order == O) {                                                          \
        if (axis == 0) return launch_hw<0, O>(e, src, dst, nx, ny, nz, taps);  \
        if (axis == 1) return launch_hw<1, O>(e, src, dst, nx, ny, nz, taps);  \
        return launch_hw<2, O>(e, src, dst, nx, ny, nz, taps);                 \
    }
    S3D_DY_ORDER(0)
    S3D_DY_ORDER(1)
    S3D_DY_ORDER(2)
#undef S3D_DY_ORDER
    return 1;
}

// z pass over the buffer planes [zb, ze) of a Z-slab buffer (see k_conv_dyadic_zr); order 1 or 2;
// returns 1 if not instantiated, -1 on a launch error
int s3d_conv_dyadic_zrange(s3d_engine *e, int order, const float *src, float *dst, int nx, int ny,
                           int nbuf, int zb, int ze, int gbase, int nglob, const TapSet &taps)
{
    if (order == 1) return launch_zr_hw<1>(e, src, dst, nx, ny, nbuf, zb, ze, gbase, nglob, taps);
    if (order == 2) return launch_zr_hw<2>(e, src, dst, nx, ny, nbuf, zb, ze, gbase, nglob, taps);
    return 1;
}
