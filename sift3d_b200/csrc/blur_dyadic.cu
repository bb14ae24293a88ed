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

template <int O, int HW, int RUN>
__host__ __device__ __forceinline__ void conv_run(const float *__restrict__ line, size_t st, int n, int i0,
                                         const TapSet &taps, float (&acc)[RUN])
{
    constexpr int P = 1 << O;
    constexpr int UHW = (HW + P - 1) >> O;  // ceil(HW / P): reach towards lower coordinates
    constexpr int OFFMAX = HW >> O;         // floor(HW / P): reach towards higher coordinates
#pragma unroll
    for (int k = 0; k < RUN; k++) acc[k] = 0.0f;
    const int last = n - 1;
    float hi = LDG(line + (size_t)s3d_clampi(i0 + RUN + OFFMAX, last) * st);
#pragma unroll
    for (int jj = RUN - 1 + OFFMAX; jj >= -UHW; jj--) {  // sample j = i0 + jj
        const float lo = LDG(line + (size_t)s3d_clampi(i0 + jj, last) * st);
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

// AXIS 0: thread = RUN consecutive x of one row; AXIS 1 / 2: thread = one x, RUN consecutive
// y / z (lanes are x neighbours: every load and store of a warp is one coalesced segment).
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

template <int AXIS, int O, int HW>
int launch_one(s3d_engine *e, const float *src, float *dst, int nx, int ny, int nz,
               const TapSet &taps)
{
    constexpr int RUN = AXIS == 0 ? 8 : 16;
    const int n = AXIS == 0 ? nx : (AXIS == 1 ? ny : nz);
    const size_t nruns = (size_t)(n + RUN - 1) / RUN;
    const size_t nthreads = (size_t)nx * ny * nz / (size_t)n * nruns;
    const size_t want = (nthreads + 255) / 256;
    const size_t cap = (size_t)e->num_sms * 64;
    const int grid = (int)(want < cap ? (want ? want : 1) : cap);
    k_conv_dyadic<AXIS, O, HW, RUN><<<grid, 256, 0, e->stream>>>(src, dst, nx, ny, nz, taps);
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
    if (!(hw == 2 || hw == 3 || hw == 4 || hw == 5 || hw == 6 || hw == 8)) return -1;
    if ((long long)n >= (1ll << 20)) return -1;
    if (uf == 1.0f) return 0;
    if (uf == 0.5f) return 1;
    if (uf == 0.25f) return 2;
    return -1;
}

// one axis of the separable filter; returns 1 if (axis, order, width) is not instantiated
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
