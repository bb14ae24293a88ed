// blur_fused.cu -- fused X/Y/Z separable Gaussian for sm_100a (the pyramid's hot kernel).
//
// Replaces apply_Sep_FIR_filter (imutil.c:3459-3544) = copy + [permute, convolve_sep_gen
// (imutil.c:2274-2393), permute] x 3 for the case every octave-0 pyramid level hits:
// one channel, tap spacing exactly 1 voxel on all three axes (unit / units == 1).
// HBM traffic is the algorithmic minimum, 8 B/voxel (read once, write once); the
// x- and y-filtered intermediates never leave the SM.
//
// Bit-exactness.  With integer tap spacing every sample coordinate c = i - d is an
// integer, so the reference's per-tap expression tap*((1-frac)*lo + frac*hi) reduces to
// tap*S(c) where S is the line extended by the reference's boundary rules
// (imutil.c:2365-2387):   S(j) = src[-j]                       j < 0
//                         S(j) = src[j]                        0 <= j < n-1
//                         S(j) = (1-f)*src[lo] + f*src[lo+1]   j >= n-1, c' = 2(n-1) - j - 0.1,
//                                                              lo = (int)c', f = c' - lo
// (for j < n-1 the "+ 0*hi" term adds an exact zero).  Each axis is therefore a plain FIR
// over an extended line, accumulated in the reference's order d = -hw..hw (= descending
// sample index) with separately rounded multiply and add.
//
// Structure.  A CTA owns a 64x32 (x,y) column and marches DOWN in z:
//   fill   global -> registers (two 16-byte loads per thread, issued one plane ahead, right
//          after the X phase) -> smem tile A (halo in x and y), rows pair-interleaved
//   X      each thread: 4 outputs along x for a PAIR of rows  (A -> smem B)
//   Y      each thread: its own 2 rows x 1 x-pair             (B -> registers)
//   Z      transposed-form FIR: 2hw+1 running partial sums per owned voxel live in
//          registers; each new xy-filtered plane updates all of them and retires one
//          output plane (coalesced 8-byte stores).  Descending z makes every output
//          accumulate its taps in the reference's order.
// All arithmetic is packed FFMA2 (fma.rn.f32x2): a*b+(-0) is the exactly rounded product,
// a*1+c the exactly rounded sum, two voxels per instruction -- measured 2x the scalar
// FMUL+FADD rate on B200 (tools/ubench.cu).  The tap symmetry t[a] == t[2hw-a] lets the Z
// phase share products: (hw+1) multiplies + (2hw+1) adds per voxel pair.
#include "common.cuh"

#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)

#include <algorithm>
#include <cmath>
#include <cstring>

namespace {

constexpr int TX = 64, TY = 32, NT = 512;
constexpr int MAXHW = 8;
constexpr int APITCH = 82;  // float2 per row pair of A; 8*APITCH % 128 == 16 -> conflict-free LDS.128
constexpr int BPITCH = 68;  // floats per row of B (16-byte aligned rows)

typedef unsigned long long u64;

__device__ __forceinline__ u64 pk(float lo, float hi)
{
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk(u64 v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c)
{
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

struct Consts {
    u64 nz, one, pz;  // (-0,-0), (1,1), (+0,+0)
};
// exactly rounded product / sum of two packed pairs
__device__ __forceinline__ u64 mul2(u64 a, u64 t, const Consts &k) { return fma2(a, t, k.nz); }
__device__ __forceinline__ u64 add2(u64 a, u64 b, const Consts &k) { return fma2(a, k.one, b); }

struct Seg {
    int x0, y0, za, zb;
};

struct MirrorTab {  // right-hand mirror samples j = n-1+k, k = 0..MAXHW (host-computed, f32)
    int lo[MAXHW + 1];
    float f[MAXHW + 1], omf[MAXHW + 1];
};

struct FusedParams {
    const float *src;
    float *dst;
    int nx, ny, nz;
    const Seg *segs;
    const int *seg_start;  // per CTA: [seg_start[b], seg_start[b+1])
    MirrorTab mx, my, mz;
    TapSet taps;
    // -0.0f / 1.0f / +0.0f passed at run time: with literal constants ptxas folds
    // fma(a,t,-0) -> mul and fma(p,1,c) -> add and then CONTRACTS the pair into one FFMA2
    // (observed in SASS; -fmad=false does not cover f32x2), which would break bit-exactness.
    float c_negzero, c_one, c_zero;
    int dbg_flags;   // timing experiments only: 1 = skip right-mirror stores, 2 = skip prev loads, 4 = skip y mirror
    long long *dbg;  // optional: per CTA {start clock, end clock, smid, steps}
};

enum { MODE_NORMAL = 0, MODE_STASH = 1, MODE_LERP = 2 };

// AL: row length a multiple of 4 -- every row start and tile column is 16-byte aligned: 16-byte
// loads, 8-byte stores.  !AL (e.g. the 181 x 217 x 181 example volume): the same kernel with
// element-wise predicated 4-byte loads and stores.
template <int HW, bool AL>
__global__ void __launch_bounds__(NT, 1) k_blur_fused(const FusedParams P)
{
    constexpr int W = 2 * HW + 1;
    constexpr int HWA = (HW + 3) & ~3;   // x halo rounded up so that 16-byte loads stay aligned
    constexpr int NR = TY + 2 * HW;      // rows of the A/B tiles (even)
    constexpr int NRP = NR / 2;          // row pairs
    constexpr int AW = TX + 2 * HWA;     // columns of the A tile
    constexpr int AWQ = AW / 4;          // 4-column groups per row
    constexpr int D0 = (HWA - HW) & ~1;  // even column where the X window of run 0 starts
    constexpr int SH = (HWA - HW) & 1;   // 1 if the true window starts one column later
    constexpr int LW = 2 * HW + 4 + 2 * SH;  // window length (even)
    constexpr int NXI = NRP * (TX / 4);  // X-phase items: (row pair, run of 4 outputs)
    static_assert(NRP * AWQ <= NT, "one fill item per thread");
    static_assert(NXI <= NT, "one X item per thread");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *Abuf = reinterpret_cast<float2 *>(smem_raw);              // 3 x [NRP][APITCH]
    float *Bbuf = reinterpret_cast<float *>(Abuf + 3 * NRP * APITCH);  // 3 x [NR][BPITCH]
    short *task_plane = reinterpret_cast<short *>(Bbuf + 3 * NR * BPITCH);
    unsigned char *task_mode = reinterpret_cast<unsigned char *>(task_plane + (P.nz + 2 * HW + 4));

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nx = P.nx, ny = P.ny, nz = P.nz;
    long long dbg_t0 = 0;
    int dbg_steps = 0;
    if (P.dbg && tid == 0) dbg_t0 = clock64();
    const size_t plane_stride = (size_t)nx * ny;
    Consts K;
    K.nz = pk(P.c_negzero, P.c_negzero);
    K.one = pk(P.c_one, P.c_one);
    K.pz = pk(P.c_zero, P.c_zero);

    // X-phase item of this thread: row pair xrp, run xrun (4 outputs); all 32 lanes of the
    // first NXI/32 warps are busy
    const bool x_thread = tid < NXI;
    const int xrun = tid / NRP, xrp = tid - xrun * NRP;
    const bool swap_st = (xrp & 4) != 0;  // bank-conflict-free order of the two 16-byte stores

    for (int si = P.seg_start[blockIdx.x]; si < P.seg_start[blockIdx.x + 1]; si++) {
        const Seg sg = P.segs[si];
        const int x0 = sg.x0, y0 = sg.y0;
        // ---- task list: samples j = zb-1+HW .. za-HW of the extended z line --------------
        __syncthreads();
        int ntask;
        {
            const int dim_end = nz - 1;
            const int jt = sg.zb - 1 + HW, jb = sg.za - HW;
            const int nlerp = jt >= dim_end ? jt - dim_end + 1 : 0;
            ntask = (jt - jb + 1) + (nlerp ? 1 : 0);
            for (int t = tid; t < ntask; t += NT) {
                int plane, mode;
                if (nlerp && t == 0) {
                    plane = P.mz.lo[jt - dim_end];  // = 2*dim_end - jt - 1
                    mode = MODE_STASH;
                } else {
                    const int j = jt - (t - (nlerp ? 1 : 0));
                    if (j >= dim_end) {
                        plane = P.mz.lo[j - dim_end] + 1;
                        mode = MODE_LERP | ((j - dim_end) << 2);
                    } else {
                        plane = j < 0 ? -j : j;
                        mode = MODE_NORMAL;
                    }
                }
                task_plane[t] = (short)plane;
                task_mode[t] = (unsigned char)mode;
            }
        }

        // ---- fill item: row pair frp, columns 4*fq..4*fq+3 of the A tile; only groups and rows
        //      inside the volume are loaded (mirror samples are synthesised by the readers) -----
        const int frp = tid / AWQ, fq = tid - frp * AWQ;
        const int fy = y0 - HW + 2 * frp;
        const int fx = x0 - HWA + 4 * fq;
        const bool fill_thread = tid < NRP * AWQ && (AL ? (fx >= 0 && fx + 3 <= nx - 1) : (fx + 3 >= 0 && fx <= nx - 1));
        const bool row0_ok = fill_thread && fy >= 0 && fy < ny;
        const bool row1_ok = fill_thread && fy + 1 >= 0 && fy + 1 < ny;
        const float *g0 = P.src + (ptrdiff_t)((size_t)(row0_ok ? fy : 0) * nx) + (fill_thread ? fx : 0);
        const float *g1 = P.src + (ptrdiff_t)((size_t)(row1_ok ? fy + 1 : 0) * nx) + (fill_thread ? fx : 0);
        const int fdst_off = frp * APITCH + 4 * fq;
        const bool fswap = (lane & 4) != 0;
        auto load4 = [&](const float *row, bool ok) -> float4 {
            if (AL) return ok ? __ldg(reinterpret_cast<const float4 *>(row)) : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok) {  // columns outside [0, nx-1] stay zero; the mirror patch fills the ones that are read
                if (fx >= 0 && fx <= nx - 1) r.x = __ldg(row);
                if (fx + 1 >= 0 && fx + 1 <= nx - 1) r.y = __ldg(row + 1);
                if (fx + 2 >= 0 && fx + 2 <= nx - 1) r.z = __ldg(row + 2);
                if (fx + 3 >= 0 && fx + 3 <= nx - 1) r.w = __ldg(row + 3);
            }
            return r;
        };
        auto store_item = [&](float2 *Adst, const float4 a, const float4 b) {
            if (!fill_thread) return;
            const float4 lo4 = make_float4(a.x, b.x, a.y, b.y), hi4 = make_float4(a.z, b.z, a.w, b.w);
            float4 *d = reinterpret_cast<float4 *>(Adst + fdst_off);
            if (fswap) {
                d[1] = hi4;
                d[0] = lo4;
            } else {
                d[0] = lo4;
                d[1] = hi4;
            }
        };

        // ---- X phase: (row pair, 4 outputs) per thread, window streamed right to left ------------
        auto xphase = [&](const float2 *Asrc, float *Bdst) {
            if (!x_thread) return;
            const float2 *arow = Asrc + xrp * APITCH;
            u64 o0 = K.pz, o1 = K.pz, o2 = K.pz, o3 = K.pz;
            auto tapstep = [&](int p, u64 v) {  // window position p, descending
                const int a0 = SH + 2 * HW - p;  // output i uses tap index a0 + i
                if (a0 >= 0 && a0 < W) o0 = add2(mul2(v, pk(P.taps.t[a0 < 0 || a0 >= W ? 0 : a0], P.taps.t[a0 < 0 || a0 >= W ? 0 : a0]), K), o0, K);
                if (a0 + 1 >= 0 && a0 + 1 < W) o1 = add2(mul2(v, pk(P.taps.t[a0 + 1 < 0 || a0 + 1 >= W ? 0 : a0 + 1], P.taps.t[a0 + 1 < 0 || a0 + 1 >= W ? 0 : a0 + 1]), K), o1, K);
                if (a0 + 2 >= 0 && a0 + 2 < W) o2 = add2(mul2(v, pk(P.taps.t[a0 + 2 < 0 || a0 + 2 >= W ? 0 : a0 + 2], P.taps.t[a0 + 2 < 0 || a0 + 2 >= W ? 0 : a0 + 2]), K), o2, K);
                if (a0 + 3 >= 0 && a0 + 3 < W) o3 = add2(mul2(v, pk(P.taps.t[a0 + 3 < 0 || a0 + 3 >= W ? 0 : a0 + 3], P.taps.t[a0 + 3 < 0 || a0 + 3 >= W ? 0 : a0 + 3]), K), o3, K);
            };
            {
                const float2 *win = arow + 4 * xrun + D0;
#pragma unroll
                for (int kk = 0; kk < LW / 2; kk++) {
                    const ulonglong2 v2 = *reinterpret_cast<const ulonglong2 *>(win + (LW - 2 - 2 * kk));
                    tapstep(LW - 1 - 2 * kk, v2.y);
                    tapstep(LW - 2 - 2 * kk, v2.x);
                }
            }
            float a0f, b0f, a1f, b1f, a2f, b2f, a3f, b3f;
            upk(o0, a0f, b0f);
            upk(o1, a1f, b1f);
            upk(o2, a2f, b2f);
            upk(o3, a3f, b3f);
            float *brow = Bdst + (2 * xrp) * BPITCH + 4 * xrun;
            const float4 r0 = make_float4(a0f, a1f, a2f, a3f), r1 = make_float4(b0f, b1f, b2f, b3f);
            if (swap_st) {
                *reinterpret_cast<float4 *>(brow + BPITCH) = r1;
                *reinterpret_cast<float4 *>(brow) = r0;
            } else {
                *reinterpret_cast<float4 *>(brow) = r0;
                *reinterpret_cast<float4 *>(brow + BPITCH) = r1;
            }
        };

        // ---- Y phase: rows y = 2*warp, 2*warp+1 of the tile; x pair = lane ---------------------
        auto yphase = [&](const float *Bsrc, u64 &y0acc, u64 &y1acc) {
            y0acc = K.pz;
            y1acc = K.pz;
            const float *top = Bsrc + 2 * lane + (2 * warp + 1 + 2 * HW) * BPITCH;
#pragma unroll
            for (int k = 0; k < W + 1; k++) {  // rows descending from 2*warp+1+HW
                const u64 v = *reinterpret_cast<const u64 *>(top - k * BPITCH);
                if (k < W) y1acc = add2(mul2(v, pk(P.taps.t[k < W ? k : 0], P.taps.t[k < W ? k : 0]), K), y1acc, K);
                if (k >= 1) y0acc = add2(mul2(v, pk(P.taps.t[k >= 1 ? k - 1 : 0], P.taps.t[k >= 1 ? k - 1 : 0]), K), y0acc, K);
            }
        };

        // ---- mirror patches (block-uniform; run one step before their consumer, so they need
        //      no barrier of their own and sit off the critical path) ----------------------------
        const bool xedge = (x0 - HW < 0) || (x0 + TX + HW > nx - 1);
        const bool yedge = (y0 - HW < 0) || (y0 + TY + HW > ny - 1);
        auto xpatch = [&](float2 *Ap) {  // columns of A outside [0, nx-1)
            const int nl = x0 - HW < 0 ? HW - x0 : 0;
            const int nrt = x0 + TX + HW > nx - 1 ? x0 + TX + HW - (nx - 1) : 0;
            const int ncol = nl + nrt;
            float *Af = reinterpret_cast<float *>(Ap);
            for (int e = tid; e < NR * ncol; e += NT) {
                const int r = e / ncol, c = e - r * ncol;
                const int base = ((r >> 1) * APITCH) * 2 + (r & 1);
                if (c < nl) {
                    const int x = x0 - HW + c;  // < 0: copy of column -x
                    Af[base + 2 * (x - (x0 - HWA))] = Af[base + 2 * (-x - (x0 - HWA))];
                } else {
                    const int k = c - nl;  // x = nx-1+k: omf*col[lo] + f*col[lo+1]
                    const int lo = P.mx.lo[k] - (x0 - HWA);
                    const float v = __fadd_rn(__fmul_rn(P.mx.omf[k], Af[base + 2 * lo]),
                                              __fmul_rn(P.mx.f[k], Af[base + 2 * (lo + 1)]));
                    Af[base + 2 * (nx - 1 + k - (x0 - HWA))] = v;  // k = 0 overwrites its own `hi`
                }
            }
        };
        auto ypatch = [&](float *Bp) {  // rows of B outside [0, ny-1)
            const int nt_ = y0 - HW < 0 ? HW - y0 : 0;
            const int nb_ = y0 + TY + HW > ny - 1 ? y0 + TY + HW - (ny - 1) : 0;
            for (int e = tid; e < (nt_ + nb_) * TX; e += NT) {
                const int rr = e / TX, x = e - rr * TX;
                if (rr < nt_) {
                    const int y = y0 - HW + rr;  // < 0: copy of row -y
                    Bp[rr * BPITCH + x] = Bp[(rr - 2 * y) * BPITCH + x];
                } else {
                    const int k = rr - nt_;  // y = ny-1+k
                    const int rl = P.my.lo[k] - (y0 - HW);
                    const float v = __fadd_rn(__fmul_rn(P.my.omf[k], Bp[rl * BPITCH + x]),
                                              __fmul_rn(P.my.f[k], Bp[(rl + 1) * BPITCH + x]));
                    Bp[(ny - 1 + k - (y0 - HW)) * BPITCH + x] = v;
                }
            }
        };

        u64 acc[2][W];  // Z-phase partial sums by age, for the thread's two rows
        u64 prevY[2];
#pragma unroll
        for (int a = 0; a < W; a++) acc[0][a] = acc[1][a] = K.pz;
        prevY[0] = prevY[1] = K.pz;
        __syncthreads();  // task table visible; previous segment's smem traffic finished

        // ---- software pipeline, ONE barrier per step.  In step t (t = -4 .. ntask-1):
        //        load  plane t+4 -> registers (stored to A[(t+4)%3] at the end of the step)
        //        patch x mirrors of A(t+3)          (stored at the end of step t-1)
        //        X     A(t+2) -> B(t+2)             (patched during step t-1)
        //        patch y mirrors of B(t+1)          (produced during step t-1)
        //        Y, Z  B(t)                         (patched during step t-1)
        //      every read is of data completed before this step's barrier; writers and readers of a
        //      step touch different ring slots, so warps drift freely between barriers.
        float *optr = P.dst + ((size_t)(sg.zb + 2 * HW) * ny + (y0 + 2 * warp)) * nx + x0 + 2 * lane;
        int zout = sg.zb + 2 * HW;
        dbg_steps += ntask + 4;
        // narrow filters do too little work per step to hide a DRAM round trip inside ONE step:
        // with PF2 the loads are issued a step earlier and parked in a second register set
        constexpr bool PF2 = HW <= 3;  // measured: helps w=5,7; hurts w>=9 (register pressure)
        float4 qa = make_float4(0.f, 0.f, 0.f, 0.f), qb = qa;
        for (int t = PF2 ? -5 : -4; t < ntask; t++) {
            __syncthreads();
            float4 na = make_float4(0.f, 0.f, 0.f, 0.f), nb = na;
            const bool ld = t + 4 < ntask;  // plane t+4 is stored at the end of this step
            {
                const int tl = PF2 ? t + 5 : t + 4;
                if (tl >= 0 && tl < ntask) {
                    const size_t off = (size_t)task_plane[tl] * plane_stride;
                    na = load4(g0 + off, row0_ok);
                    nb = load4(g1 + off, row1_ok);
                }
            }
            if (xedge && t + 3 >= 0 && t + 3 < ntask) xpatch(Abuf + ((t + 3) % 3) * NRP * APITCH);
            if (t + 2 >= 0 && t + 2 < ntask)
                xphase(Abuf + ((t + 2) % 3) * NRP * APITCH, Bbuf + ((t + 2) % 3) * NR * BPITCH);
            if (yedge && t + 1 >= 0 && t + 1 < ntask) ypatch(Bbuf + ((t + 1) % 3) * NR * BPITCH);
            if (t < 0) {
                if (PF2) {
                    if (ld && t + 4 >= 0) store_item(Abuf + ((t + 4) % 3) * NRP * APITCH, qa, qb);
                    qa = na, qb = nb;
                } else if (ld) {
                    store_item(Abuf + ((t + 4) % 3) * NRP * APITCH, na, nb);
                }
                continue;
            }
            u64 y0acc, y1acc;
            yphase(Bbuf + (t % 3) * NR * BPITCH, y0acc, y1acc);

            // ---- Z phase --------------------------------------------------------------------
            const int mode = task_mode[t];
            if ((mode & 3) == MODE_STASH) {
                prevY[0] = y0acc;
                prevY[1] = y1acc;
            } else {
                u64 zin[2] = {y0acc, y1acc};
                if ((mode & 3) == MODE_LERP) {
                    const int k = mode >> 2;
                    const u64 f = pk(P.mz.f[k], P.mz.f[k]), omf = pk(P.mz.omf[k], P.mz.omf[k]);
#pragma unroll
                    for (int r = 0; r < 2; r++) {
                        const u64 cur = zin[r];
                        zin[r] = add2(mul2(prevY[r], omf, K), mul2(cur, f, K), K);
                        prevY[r] = cur;
                    }
                }
                zout--;  // output z = j + HW completes with this sample
                optr -= plane_stride;
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    u64 prod[HW + 1];
#pragma unroll
                    for (int m = 0; m <= HW; m++)
                        prod[m] = mul2(zin[r], pk(P.taps.t[m], P.taps.t[m]), K);
#pragma unroll
                    for (int a = W - 1; a >= 1; a--)
                        acc[r][a] = add2(prod[a <= HW ? a : 2 * HW - a], acc[r][a - 1], K);
                    acc[r][0] = add2(prod[0], K.pz, K);
                }
                if (zout < sg.zb && zout >= sg.za) {
                    if (AL) {
                        *reinterpret_cast<u64 *>(optr) = acc[0][W - 1];
                        *reinterpret_cast<u64 *>(optr + nx) = acc[1][W - 1];
                    } else {
                        float a0, a1, b0, b1;
                        upk(acc[0][W - 1], a0, a1);
                        upk(acc[1][W - 1], b0, b1);
                        optr[0] = a0, optr[1] = a1;
                        optr[nx] = b0, optr[nx + 1] = b1;
                    }
                }
            }
            if (PF2) {
                if (ld) store_item(Abuf + ((t + 4) % 3) * NRP * APITCH, qa, qb);
                qa = na, qb = nb;
            } else if (ld) {
                store_item(Abuf + ((t + 4) % 3) * NRP * APITCH, na, nb);
            }
        }
    }
    if (P.dbg && tid == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        P.dbg[4 * blockIdx.x + 0] = dbg_t0;
        P.dbg[4 * blockIdx.x + 1] = clock64();
        P.dbg[4 * blockIdx.x + 2] = smid;
        P.dbg[4 * blockIdx.x + 3] = dbg_steps;
    }
}

// =====================================================================================================
// k_blur_tma -- the same fused blur, second design (round 2).  What changed, and why (ncu on
// k_blur_fused: 56 warp-instructions per 32 voxels at w = 5 of which 15 are FFMA2; the rest is the
// fill path, transposition MOVs, the rotation of the Z partial sums and per-step bookkeeping
// amortised over only 4 voxels per thread):
//   fill   TMA (cp.async.bulk.tensor.3d, out-of-bounds elements zero-filled) + one mbarrier per ring
//          slot, issued by one thread NSA+1 planes ahead: no LDG / STS / address arithmetic in the
//          compute warps, and the tile lies in shared memory as it lies in global memory (TMA cannot
//          interleave two rows, which is what k_blur_fused's row-pair FFMA2 operands need).
//   X      item = R consecutive outputs of ONE row (R = 8, or 16 with the tall tile).  The packed
//          operand is the TAP pair (t[a], t[a+1]) and the sample is the broadcast scalar:
//          (o[x], o[x+1]) += (t[a], t[a+1]) * in[s] with a = x - s + hw.  Every sample is used as a
//          scalar, so no operand needs a particular register-pair alignment (a sliding window of
//          x-PAIRS would need each pair at both alignments); the tap pairs at both alignments are two
//          small tables.  Each output still receives its taps in the reference's order (samples
//          descending); the table is padded with t[-1] = t[w] = +0, which adds an exact zero before
//          the first / after the last tap of one lane -- a no-op, because a partial sum that started
//          from +0 is never -0.  (Only non-finite input differs: 0 * Inf = NaN where the reference
//          has Inf.)  X results are x-pairs: no transposition before the Y phase.
//   Y, Z   a thread owns RPT rows x one x-pair (RPT = 4 for the narrow filters: a 64 x 64 tile,
//          8 voxels per thread and step; RPT = 2 where the 2 x RPT x w partial sums would not fit)
//   loop   unrolled three times: the rotation of the partial sums costs register moves only at
//          the back edge (every third plane)
// Arithmetic and boundary rules are those of k_blur_fused (bit-identical output; the parity tests
// run both against the oracle).  TMA constraints found on B200 (tools/tma_probe.cu): the box must
// fit into the tensor in every dimension and its x origin must be 16-byte aligned -- so the x
// halo is rounded up to a multiple of 4 columns and small volumes stay with k_blur_fused.
// =====================================================================================================
constexpr int BP2 = 66;  // floats per row of B: 8-byte aligned rows, lanes walking down the rows hit different banks (STS.64)

__host__ __device__ constexpr int tma_hwa(int hw) { return (hw + 3) & ~3; }
__host__ __device__ constexpr int tma_aw(int hw)
{  // box width: TX + 2 x halo, with an ODD number of 16-byte chunks per row, so that lanes walking
   // down the rows hit different banks (LDS.128)
    int aw = TX + 2 * tma_hwa(hw);
    if (((aw / 4) & 1) == 0) aw += 4;
    return aw;
}
__host__ __device__ constexpr int tma_slot_bytes(int hw, int rpt)
{
    return (((16 * rpt + 2 * hw) * tma_aw(hw) * 4) + 127) & ~127;
}
__host__ __device__ constexpr int tma_b_bytes(int hw, int rpt) { return 3 * (16 * rpt + 2 * hw) * BP2 * 4; }
constexpr int TMA_NSA = 4;  // A-ring slots

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MBAR_DONE;\n"
        "bra MBAR_WAIT;\n"
        "MBAR_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap *map, int x, int y, int z, unsigned bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
        "l"(map), "r"(x), "r"(y), "r"(z), "r"(bar)
        : "memory");
}

struct TapPairs {  // (t[a], t[a+1]) for even a (e[a/2]) and odd a (o[(a+1)/2]), t[-1] = t[w] = +0
    float2 e[MAXHW + 1], o[MAXHW + 1];
};

template <int HW, int RPT>
__global__ void __launch_bounds__(NT, 1)
    k_blur_tma(const __grid_constant__ CUtensorMap tmap, const FusedParams P, const TapPairs TP)
{
    constexpr int W = 2 * HW + 1;
    constexpr int HWA = tma_hwa(HW);     // x halo fetched (column 0 of A is x0 - HWA)
    constexpr int OFF = HWA - HW;        // first column of run 0's window
    constexpr int TYT = 16 * RPT;        // tile rows
    constexpr int NR = TYT + 2 * HW;     // rows of the A / B tiles
    constexpr int AW = tma_aw(HW);       // columns of A
    constexpr int R = RPT == 4 ? 16 : 8;  // outputs of an X item
    constexpr int L = R + 2 * HW;        // its window
    constexpr int NQ = (OFF + L + 3) / 4;  // 16-byte loads covering it
    constexpr int NXI = NR * (TX / R);   // X items of a plane
    constexpr int NSA = TMA_NSA, PDA = NSA + 1;
    constexpr int SLOTB = tma_slot_bytes(HW, RPT);
    constexpr unsigned TXB = (unsigned)NR * AW * 4u;  // bytes one plane's box delivers
    static_assert(NXI <= NT, "one X item per thread");
    extern __shared__ unsigned char smem_dyn[];
    // 128-byte aligned window (TMA destinations)
    unsigned char *smem_raw = smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u);
    float *Abuf = reinterpret_cast<float *>(smem_raw);                    // NSA x [NR][AW]
    float *Bbuf = reinterpret_cast<float *>(smem_raw + NSA * SLOTB);       // 3 x [NR][BP2]
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(Bbuf + 3 * NR * BP2);
    short *task_plane = reinterpret_cast<short *>(bars + NSA);
    unsigned char *task_mode = reinterpret_cast<unsigned char *>(task_plane + (P.nz + 2 * HW + 4));

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nx = P.nx, ny = P.ny, nz = P.nz;
    const size_t plane_stride = (size_t)nx * ny;
    Consts K;
    K.nz = pk(P.c_negzero, P.c_negzero);
    K.one = pk(P.c_one, P.c_one);
    K.pz = pk(P.c_zero, P.c_zero);
    const unsigned bar0 = smem_u32(bars), a0s = smem_u32(Abuf);

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < NSA; i++) mbar_init(bar0 + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    unsigned segbase = 0;  // planes fetched by this CTA so far: plane G sits in slot G % NSA, phase G / NSA

    // X item of this thread: row xrow of the tile (lanes = consecutive rows), outputs R*xrun .. +R-1
    const bool x_thread = tid < NXI;
    const int xrun = tid / NR, xrow = tid - xrun * NR;
    const bool producer = tid == NT - 32;  // lane 0 of the last warp (never an X warp)

    for (int si = P.seg_start[blockIdx.x]; si < P.seg_start[blockIdx.x + 1]; si++) {
        const Seg sg = P.segs[si];
        const int x0 = sg.x0, y0 = sg.y0;
        // ---- task list: samples j = zb-1+HW .. za-HW of the extended z line (as in k_blur_fused) ----
        __syncthreads();
        int ntask, tmain;
        {
            const int dim_end = nz - 1;
            const int jt = sg.zb - 1 + HW, jb = sg.za - HW;
            const int nlerp = jt >= dim_end ? jt - dim_end + 1 : 0;
            ntask = (jt - jb + 1) + (nlerp ? 1 : 0);
            tmain = nlerp ? nlerp + 1 : 0;  // the stashed plane and the lerped samples come first
            for (int t = tid; t < ntask; t += NT) {
                int plane, mode;
                if (nlerp && t == 0) {
                    plane = P.mz.lo[jt - dim_end];
                    mode = MODE_STASH;
                } else {
                    const int j = jt - (t - (nlerp ? 1 : 0));
                    if (j >= dim_end) {
                        plane = P.mz.lo[j - dim_end] + 1;
                        mode = MODE_LERP | ((j - dim_end) << 2);
                    } else {
                        plane = j < 0 ? -j : j;
                        mode = MODE_NORMAL;
                    }
                }
                task_plane[t] = (short)plane;
                task_mode[t] = (unsigned char)mode;
            }
        }
        const bool xedge = (x0 - HW < 0) || (x0 + TX + HW > nx - 1);
        const bool yedge = (y0 - HW < 0) || (y0 + TYT + HW > ny - 1);
        const int xo = x0 - HWA;  // global x of column 0 of A (a multiple of 4: 16-byte aligned origin)

        // ---- producer: plane p of the task list -> ring slot, one box, one mbarrier phase ---------
        auto produce = [&](int p) {
            if (p < 0 || p >= ntask) return;
            const unsigned G = segbase + (unsigned)p, slot = G % NSA;
            const unsigned bar = bar0 + 8 * slot;
            const int z = task_plane[p];
            // the slot's previous contents were read (and, in edge tiles, patched) through the generic
            // proxy before the barrier this thread has just passed
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(bar, TXB);
            tma_load_3d(a0s + slot * SLOTB, &tmap, xo, y0 - HW, z, bar);
        };
        auto wait_plane = [&](int p) {
            const unsigned G = segbase + (unsigned)p;
            mbar_wait(bar0 + 8 * (G % NSA), (G / NSA) & 1u);
        };
        auto a_slot = [&](int p) -> float * {
            const unsigned G = segbase + (unsigned)p;
            return Abuf + (size_t)(G % NSA) * (SLOTB / 4);
        };

        // ---- X phase: R outputs of one row per thread, samples streamed right to left.  Window
        //      column c (sample s = R*xrun - HW + c) meets output pair q with the tap pair
        //      (t[a], t[a+1]), a = 2q + 2HW - c in [-1, W-1]. -------------------------------------------
        auto xphase = [&](const float *Ap, float *Bdst) {
            if (!x_thread) return;
            const float *er = Ap + xrow * AW + R * xrun;
            float e[4 * NQ];
#pragma unroll
            for (int i = NQ - 1; i >= 0; i--) {
                const float4 v = *reinterpret_cast<const float4 *>(er + 4 * i);
                e[4 * i] = v.x, e[4 * i + 1] = v.y, e[4 * i + 2] = v.z, e[4 * i + 3] = v.w;
            }
            u64 o[R / 2];
#pragma unroll
            for (int q = 0; q < R / 2; q++) o[q] = K.pz;
#pragma unroll
            for (int c = L - 1; c >= 0; c--) {
                const u64 v = pk(e[OFF + c], e[OFF + c]);
#pragma unroll
                for (int q = 0; q < R / 2; q++) {
                    const int a = 2 * q + 2 * HW - c;
                    if (a >= -1 && a <= W - 1) {
                        const int ai = a < -1 || a > W - 1 ? 0 : a;
                        const float2 tp = (ai & 1) ? TP.o[(ai + 1) / 2] : TP.e[ai / 2];
                        o[q] = add2(mul2(pk(tp.x, tp.y), v, K), o[q], K);
                    }
                }
            }
            // 8-byte stores: ptxas does not place two FFMA2 results in one aligned register quad (a
            // 16-byte store costs four moves); with BP2 = 2 (mod 4) they are conflict-free
            u64 *brow = reinterpret_cast<u64 *>(Bdst + xrow * BP2 + R * xrun);
#pragma unroll
            for (int q = 0; q < R / 2; q++) brow[q] = o[q];
        };

        // ---- Y phase: rows RPT*warp .. +RPT-1 of the tile, x pair = lane; B rows streamed top-down
        auto yphase = [&](const float *Bsrc, u64 *yacc) {
#pragma unroll
            for (int j = 0; j < RPT; j++) yacc[j] = K.pz;
            const float *top = Bsrc + 2 * lane + (RPT * warp + RPT - 1 + 2 * HW) * BP2;
#pragma unroll
            for (int k = 0; k < W + RPT - 1; k++) {
                const u64 v = *reinterpret_cast<const u64 *>(top - k * BP2);
#pragma unroll
                for (int j = RPT - 1; j >= 0; j--) {
                    const int a = k - (RPT - 1 - j);
                    if (a >= 0 && a < W) yacc[j] = add2(mul2(v, pk(P.taps.t[a < 0 || a >= W ? 0 : a], P.taps.t[a < 0 || a >= W ? 0 : a]), K), yacc[j], K);
                }
            }
        };

        // ---- mirror patches (edge tiles only; each runs a step before its consumer).  The warps
        //      without X items do them: they have half the work of the others in every step, so
        //      the patches stay off the critical path of the step ---------------------------------------
        constexpr int PT0 = ((NXI + 31) / 32) * 32, PTN = NT - PT0;  // first patch thread, their number
        static_assert(PTN >= 64, "at least two warps without X items");
        const bool patch_thread = tid >= PT0;
        auto xpatch = [&](float *Ep) {  // columns outside [0, nx-1)
            const int nl = x0 - HW < 0 ? HW - x0 : 0;  // x = x0-HW .. -1
            const int nrt = x0 + TX + HW > nx - 1 ? x0 + TX + HW - (nx - 1) : 0;  // x = nx-1 ..
            const int ncol = nl + nrt;
            for (int e = tid - PT0; e < NR * ncol; e += PTN) {
                const int r = e / ncol, c = e - r * ncol;
                float v;
                int dc;
                if (c < nl) {
                    const int x = x0 - HW + c;  // < 0: copy of column -x
                    dc = x - xo;
                    v = Ep[r * AW + (-x - xo)];
                } else {
                    const int k = c - nl;  // x = nx-1+k: omf*col[lo] + f*col[lo+1]
                    const int lo = P.mx.lo[k] - xo;
                    v = __fadd_rn(__fmul_rn(P.mx.omf[k], Ep[r * AW + lo]), __fmul_rn(P.mx.f[k], Ep[r * AW + lo + 1]));
                    dc = nx - 1 + k - xo;  // k = 0 overwrites its own `hi` (no other entry reads it)
                }
                Ep[r * AW + dc] = v;
            }
        };
        auto ypatch = [&](float *Bp) {  // rows of B outside [0, ny-1)
            const int nt_ = y0 - HW < 0 ? HW - y0 : 0;
            const int nb_ = y0 + TYT + HW > ny - 1 ? y0 + TYT + HW - (ny - 1) : 0;
            for (int e = tid - PT0; e < (nt_ + nb_) * TX; e += PTN) {
                const int rr = e / TX, x = e - rr * TX;
                if (rr < nt_) {
                    const int y = y0 - HW + rr;  // < 0: copy of row -y
                    Bp[rr * BP2 + x] = Bp[(rr - 2 * y) * BP2 + x];
                } else {
                    const int k = rr - nt_;  // y = ny-1+k
                    const int rl = P.my.lo[k] - (y0 - HW);
                    const float v = __fadd_rn(__fmul_rn(P.my.omf[k], Bp[rl * BP2 + x]),
                                              __fmul_rn(P.my.f[k], Bp[(rl + 1) * BP2 + x]));
                    Bp[(ny - 1 + k - (y0 - HW)) * BP2 + x] = v;
                }
            }
        };

        u64 acc[RPT][W];  // Z-phase partial sums by age
        u64 prevY[RPT];
#pragma unroll
        for (int j = 0; j < RPT; j++) {
#pragma unroll
            for (int a = 0; a < W; a++) acc[j][a] = K.pz;
            prevY[j] = K.pz;
        }
        float *op[RPT];  // output pointers of the thread's rows, one plane above the next output
#pragma unroll
        for (int j = 0; j < RPT; j++)
            op[j] = P.dst + ((size_t)(sg.zb + 2 * HW) * ny + (y0 + RPT * warp + j)) * nx + x0 + 2 * lane;
        int zout = sg.zb + 2 * HW;
        // B ring: plane p sits in slot (p - tmain) mod 3, so that the unrolled main loop sees
        // compile-time slots
        float *const B0 = Bbuf, *const B1 = Bbuf + NR * BP2, *const B2 = Bbuf + 2 * NR * BP2;
        auto b_slot = [&](int p) -> float * { return Bbuf + ((p - tmain + 12) % 3) * (NR * BP2); };

        // Z phase: the xy-filtered plane `zin` updates every partial sum and retires one output plane
        auto zupdate = [&](const u64 *zin) {
            zout--;  // output z = j + HW completes with this sample
#pragma unroll
            for (int j = 0; j < RPT; j++) {
                op[j] -= plane_stride;
                u64 prod[HW + 1];
#pragma unroll
                for (int m = 0; m <= HW; m++) prod[m] = mul2(zin[j], pk(P.taps.t[m], P.taps.t[m]), K);
#pragma unroll
                for (int a = W - 1; a >= 1; a--) acc[j][a] = add2(prod[a <= HW ? a : 2 * HW - a], acc[j][a - 1], K);
                acc[j][0] = add2(prod[0], K.pz, K);
            }
            if (zout < sg.zb && zout >= sg.za) {
#pragma unroll
                for (int j = 0; j < RPT; j++) *reinterpret_cast<u64 *>(op[j]) = acc[j][W - 1];
            }
        };

        // ---- software pipeline, ONE block barrier per step.  In step t:
        //        TMA   plane t+PDA -> A ring                 (slot last read in step t-1)
        //        patch x mirrors of A(t+3)                   (edge tiles; waits for its mbarrier)
        //        Y, Z  B(t)                                  (patched in step t-1)
        //        patch y mirrors of B(t+1)                   (produced in step t-1)
        //        X     A(t+2) -> B(t+2)                      (waits for the plane's mbarrier)
        //      every generic-proxy read is of data completed before this step's barrier; writers
        //      and readers of a step touch different ring slots.
        // slow steps: the pipeline fill (t < 0) and the tasks of the right-hand z mirror (stash /
        // lerp: the first tmain tasks of a segment that reaches the top of the volume)
        for (int t = -PDA; t < tmain; t++) {
            __syncthreads();
            if (producer) produce(t + PDA);
            if (xedge && patch_thread && t + 3 >= 0 && t + 3 < ntask) {
                wait_plane(t + 3);
                xpatch(a_slot(t + 3));
            }
            if (t >= 0) {
                u64 zin[RPT];
                yphase(b_slot(t), zin);
                const int mode = task_mode[t];
                if ((mode & 3) == MODE_STASH) {
#pragma unroll
                    for (int j = 0; j < RPT; j++) prevY[j] = zin[j];
                } else {  // MODE_LERP
                    const int k = mode >> 2;
                    const u64 f = pk(P.mz.f[k], P.mz.f[k]), omf = pk(P.mz.omf[k], P.mz.omf[k]);
#pragma unroll
                    for (int j = 0; j < RPT; j++) {
                        const u64 cur = zin[j];
                        zin[j] = add2(mul2(prevY[j], omf, K), mul2(cur, f, K), K);
                        prevY[j] = cur;
                    }
                    zupdate(zin);
                }
            }
            if (yedge && patch_thread && t + 1 >= 0 && t + 1 < ntask) ypatch(b_slot(t + 1));
            if (t + 2 >= 0 && t + 2 < ntask) {
                if (x_thread) wait_plane(t + 2);
                xphase(a_slot(t + 2), b_slot(t + 2));
            }
        }
        auto step = [&](int t, float *Bt, float *Bt1, float *Bt2) {
            __syncthreads();
            if (producer) produce(t + PDA);
            if (xedge && patch_thread && t + 3 < ntask) {
                wait_plane(t + 3);
                xpatch(a_slot(t + 3));
            }
            u64 zin[RPT];
            yphase(Bt, zin);
            zupdate(zin);
            if (yedge && patch_thread && t + 1 < ntask) ypatch(Bt1);
            if (t + 2 < ntask) {
                if (x_thread) wait_plane(t + 2);
                xphase(a_slot(t + 2), Bt2);
            }
        };
        for (int t = tmain; t < ntask;) {
            step(t, B0, B1, B2);
            if (++t >= ntask) break;
            step(t, B1, B2, B0);
            if (++t >= ntask) break;
            step(t, B2, B0, B1);
            if (++t >= ntask) break;
        }
        segbase += (unsigned)ntask;
    }
}

template <int RPT>
size_t tma_smem_bytes(int hw, int nz)
{
    size_t b = (size_t)TMA_NSA * tma_slot_bytes(hw, RPT) + tma_b_bytes(hw, RPT) + TMA_NSA * 8;
    b += (size_t)(nz + 2 * hw + 4) * sizeof(short) + (size_t)(nz + 2 * hw + 4);
    return ((b + 15) & ~(size_t)15) + 128;  // + alignment slack
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn tma_encoder()
{
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

template <int HW, int RPT>
int launch_tma(s3d_engine *e, const FusedParams &P, int grid, int nzbuf)
{
    CUtensorMap map;
    const cuuint64_t dims[3] = {(cuuint64_t)P.nx, (cuuint64_t)P.ny, (cuuint64_t)nzbuf};
    const cuuint64_t strides[2] = {(cuuint64_t)P.nx * 4, (cuuint64_t)P.nx * P.ny * 4};
    const cuuint32_t box[3] = {(cuuint32_t)tma_aw(HW), (cuuint32_t)(16 * RPT + 2 * HW), 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult cr = tma_encoder()(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(P.src), dims, strides,
                                      box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return s3d_fail(e, "fused blur: cuTensorMapEncodeTiled", cudaErrorInvalidValue, __FILE__, __LINE__);
    TapPairs TP;
    auto tap = [&](int a) { return a >= 0 && a < P.taps.width ? P.taps.t[a] : 0.0f; };
    for (int k = 0; k <= MAXHW; k++) {
        TP.e[k] = make_float2(tap(2 * k), tap(2 * k + 1));
        TP.o[k] = make_float2(tap(2 * k - 1), tap(2 * k));
    }
    const size_t smem = tma_smem_bytes<RPT>(HW, P.nz);
    static bool attr_set[64] = {};
    if (!attr_set[e->device & 63]) {
        S3D_CUDA(e, cudaFuncSetAttribute(k_blur_tma<HW, RPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set[e->device & 63] = true;
    }
    k_blur_tma<HW, RPT><<<grid, NT, smem, e->stream>>>(map, P, TP);
    S3D_LAUNCH_CHECK(e);
    return 0;
}

size_t smem_bytes(int hw, int nz)
{
    const int NR = TY + 2 * hw;
    size_t b = (size_t)3 * (NR / 2) * APITCH * sizeof(float2) + (size_t)3 * NR * BPITCH * sizeof(float);
    b += (size_t)(nz + 2 * hw + 4) * sizeof(short) + (size_t)(nz + 2 * hw + 4);
    return (b + 15) & ~(size_t)15;
}

void mirror_table(int n, MirrorTab &m)
{  // literal f32 evaluation of the right-hand mirror (imutil.c:2376-2380)
    const int dim_end = n - 1;
    for (int k = 0; k <= MAXHW; k++) {
        volatile float c = (float)(dim_end + k);
        volatile float t = 2.0f * (float)dim_end;
        t = t - c;
        t = t - 0.1f;
        const int lo = (int)t;
        volatile float f = t - (float)lo;
        volatile float omf = 1.0f - f;
        m.lo[k] = lo;
        m.f[k] = f;
        m.omf[k] = omf;
    }
}

template <int HW, bool AL>
int launch_al(s3d_engine *e, const FusedParams &P, int grid, size_t smem)
{
    static bool attr_set[64] = {};
    if (!attr_set[e->device & 63]) {
        S3D_CUDA(e, cudaFuncSetAttribute(k_blur_fused<HW, AL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         200 * 1024));
        attr_set[e->device & 63] = true;
    }
    k_blur_fused<HW, AL><<<grid, NT, smem, e->stream>>>(P);
    S3D_LAUNCH_CHECK(e);
    return 0;
}

template <int HW>
int launch(s3d_engine *e, const FusedParams &P, int grid, size_t smem)
{
    return (P.nx & 3) ? launch_al<HW, false>(e, P, grid, smem) : launch_al<HW, true>(e, P, grid, smem);
}

}  // namespace

bool s3d_blur_fused_eligible(int nx, int ny, int nz, int nc, const TapSet &taps, const float uf[3])
{
    const int hw = taps.width / 2;
    if (nc != 1 || hw < 1 || hw > MAXHW) return false;
    if (uf[0] != 1.0f || uf[1] != 1.0f || uf[2] != 1.0f) return false;
    if (nx < TX || ny < TY || nz < 2 * hw + 2 || nx < 2 * hw + 2 || ny < 2 * hw + 2) return false;
    if (nz + 2 * hw + 4 > 32000) return false;  // short task table
    for (int i = 0; i < taps.width; i++)
        if (taps.t[i] != taps.t[taps.width - 1 - i]) return false;  // Z phase shares products
    if (smem_bytes(hw, nz) > 200 * 1024) return false;
    return true;
}

int s3d_blur_fused_zrange(s3d_engine *e, const float *src, float *dst, int nx, int ny, int nz,
                          const TapSet &taps, int zb, int ze, int gz0, int nz_glob);

int s3d_blur_fused(s3d_engine *e, const float *src, float *dst, int nx, int ny, int nz,
                   const TapSet &taps, const float uf[3])
{
    (void)uf;
    return s3d_blur_fused_zrange(e, src, dst, nx, ny, nz, taps, 0, nz, 0, nz);
}

// Output planes [zb, ze) only (Z-slab tiling: the planes outside are halo, filled by the
// neighbours).  The buffer is the window [gz0, gz0 + nz) of a volume of nz_glob planes; the z
// mirror rules key on plane 0 / nz-1 of the buffer, which a tiled caller arranges to be true
// volume ends or out of the filter's reach.  The right-hand mirror weights are the f32
// roundings of expressions in the GLOBAL plane index, so the table is built from nz_glob.
int s3d_blur_fused_zrange(s3d_engine *e, const float *src, float *dst, int nx, int ny, int nz,
                          const TapSet &taps, int zb, int ze, int gz0, int nz_glob)
{
    const int hw = taps.width / 2;
    const int nzr = ze - zb;
    if (nzr <= 0) return 0;
    // which kernel: k_blur_tma (rows a multiple of 4 voxels, 16-byte aligned base -- what the tensor
    // map needs) with a 64 x 64 tile for the narrow filters, else k_blur_fused
    const bool use_tma = e->opt_blur_v1 == 0 && (nx & 3) == 0 && ((uintptr_t)src & 15) == 0 && nx >= tma_aw(hw) &&
                         ny >= 32 + 2 * hw && tma_encoder() != nullptr && tma_smem_bytes<2>(hw, nz) <= 227 * 1024;
    const int rpt = use_tma && hw <= e->opt_blur_rpt4_hw && hw <= 4 && ny >= 64 + 2 * hw &&
                            tma_smem_bytes<4>(hw, nz) <= 227 * 1024 ? 4 : 2;
    const int TYv = use_tma ? 16 * rpt : TY;
    // ---- work decomposition: columns (overlapping last tile) x balanced z ranges,
    //      computed once per volume size and cached in the engine -------------------------
    const SegTab *tab = nullptr;
    for (const auto &t : e->segtabs)
        if (t.nx == nx && t.ny == ny && t.nz == nz && t.zb == zb && t.ze == ze && t.ty == TYv && t.hw == hw) tab = &t;
    if (!tab) {
        std::vector<int> xs, ys;
        for (int x = 0; x < nx; x += TX) xs.push_back(std::min(x, nx - TX));
        for (int y = 0; y < ny; y += TYv) ys.push_back(std::min(y, ny - TYv));
        const long ncol = (long)xs.size() * ys.size();
        const long total = ncol * nzr;
        const int grid = (int)std::min<long>(e->num_sms, std::max<long>(1, total / 16));
        // Edge columns cost more per plane (mirror samples are synthesised by a few threads on the
        // critical path of every step); measured on B200 (tools/blur_dbg.py, cycles/step relative
        // to an interior column): left 1.11, right 1.30, top 1.06, bottom 1.19.  The z ranges are
        // balanced by cost so that all persistent CTAs finish together.
        // cost of an edge column relative to an interior one (left, right, top, bottom): the mirror
        // patches grow with the filter.  Measured with k_blur_tma at 512^3 (profiles/r02_blur_ab.txt):
        // the fastest of five weight sets per filter width
        static const double bw_tma[3][4] = {{1.02, 1.04, 1.02, 1.04}, {1.01, 1.02, 1.01, 1.02}, {1.10, 1.20, 1.10, 1.20}};
        const double *bw = (use_tma && !e->blur_w_user) ? bw_tma[hw <= 3 ? 0 : hw == 4 ? 1 : 2] : e->blur_w;
        std::vector<double> wcol(ncol);
        double wsum = 0;
        for (long c = 0; c < ncol; c++) {
            const int x0 = xs[c % xs.size()], y0 = ys[c / xs.size()];
            double w = 1.0;
            if (x0 - hw < 0) w *= bw[0];
            if (x0 + TX + hw > nx - 1) w *= bw[1];
            if (y0 - hw < 0) w *= bw[2];
            if (y0 + TYv + hw > ny - 1) w *= bw[3];
            wcol[c] = w;
            wsum += w * nzr;
        }
        std::vector<Seg> segs;
        std::vector<int> start(grid + 1, 0);
        {
            // The work list is cut into `grid` consecutive shares of equal cost.  Its order decides what
            // runs at the same time: CTA b starts at list position b * quota, so neighbouring columns
            // are (quota mod piece length) planes apart in z, and their shared halo columns / rows hit
            // in L2 only when that is well below the ~120 planes of 512 x 512 floats the L2 holds.
            // With one piece per column and fewer columns than CTAs (the 64 x 64 tile: 64 columns,
            // quota 221 of 512 planes) they are 221 planes apart and every halo is fetched from DRAM
            // again (measured 1.30 x the algorithmic traffic); cutting z into S slabs and listing the
            // pieces slab by slab brings the offset to |quota - nz / S| (S = 2: 35 planes).
            const int S = e->opt_blur_slabs > 0 ? e->opt_blur_slabs
                                                : (int)std::max<long>(1, std::min<long>(std::lround((double)grid / (double)ncol), nzr / (4 * hw + 4)));
            const long npiece = ncol * S;
            auto piece_lo = [&](long pc) { return zb + (int)((long)nzr * (pc / ncol) / S); };
            auto piece_hi = [&](long pc) { return zb + (int)((long)nzr * (pc / ncol + 1) / S); };
            const double quota = wsum / grid;
            long pc = 0;
            int z = piece_lo(0);
            for (int b = 0; b < grid; b++) {
                start[b] = (int)segs.size();
                double need = quota;
                while (pc < npiece && (need > 1e-9 || b == grid - 1)) {
                    const long col = pc % ncol;
                    const int pze = piece_hi(pc);
                    const double per = wcol[col];  // halo planes ignored
                    int take = (b == grid - 1) ? pze - z : (int)std::min<double>(pze - z, std::ceil(need / per - 1e-9));
                    if (take <= 0) break;
                    // avoid leaving a sliver shorter than the halo at the end of a piece
                    if (pze - (z + take) > 0 && pze - (z + take) < 2 * hw && b != grid - 1) take = pze - z;
                    Seg sg;
                    sg.x0 = xs[col % xs.size()];
                    sg.y0 = ys[col / xs.size()];
                    sg.za = z;
                    sg.zb = z + take;
                    segs.push_back(sg);
                    need -= take * per;
                    z += take;
                    if (z >= pze) {
                        pc++;
                        if (pc < npiece) z = piece_lo(pc);
                    }
                }
            }
        }
        start[grid] = (int)segs.size();
        const size_t need = segs.size() * sizeof(Seg) + start.size() * sizeof(int);
        std::vector<unsigned char> host(need);
        memcpy(host.data(), segs.data(), segs.size() * sizeof(Seg));
        memcpy(host.data() + segs.size() * sizeof(Seg), start.data(), start.size() * sizeof(int));
        SegTab nt;
        nt.nx = nx, nt.ny = ny, nt.nz = nz, nt.grid = grid, nt.nseg = segs.size(), nt.d = nullptr;
        nt.zb = zb, nt.ze = ze, nt.ty = TYv, nt.hw = hw;
        S3D_CUDA(e, cudaMalloc(&nt.d, need));
        S3D_CUDA(e, cudaMemcpyAsync(nt.d, host.data(), need, cudaMemcpyHostToDevice, e->stream));
        S3D_CUDA(e, cudaStreamSynchronize(e->stream));
        e->segtabs.push_back(nt);
        tab = &e->segtabs.back();
    }
    const int grid = tab->grid;
    const void *d_tab = tab->d;
    const size_t nseg = tab->nseg;

    FusedParams P;
    P.src = src;
    P.dst = dst;
    P.nx = nx;
    P.ny = ny;
    P.nz = nz;
    P.segs = reinterpret_cast<const Seg *>(d_tab);
    P.seg_start = reinterpret_cast<const int *>((const unsigned char *)d_tab + nseg * sizeof(Seg));
    mirror_table(nx, P.mx);
    mirror_table(ny, P.my);
    mirror_table(nz_glob, P.mz);
    for (int k = 0; k <= MAXHW; k++) P.mz.lo[k] -= gz0;  // global plane -> buffer plane
    P.taps = taps;
    P.c_negzero = -0.0f;
    P.c_one = 1.0f;
    P.c_zero = 0.0f;
    P.dbg = e->d_blur_dbg;
    P.dbg_flags = e->opt_blur_flags;
    if (use_tma) {
#define S3D_TMA_CASE(H) \
    case H: return rpt == 4 ? launch_tma<H, 4>(e, P, grid, nz) : launch_tma<H, 2>(e, P, grid, nz);
        switch (hw) {
            S3D_TMA_CASE(1)
            S3D_TMA_CASE(2)
            S3D_TMA_CASE(3)
            S3D_TMA_CASE(4)
        case 5: return launch_tma<5, 2>(e, P, grid, nz);
        case 6: return launch_tma<6, 2>(e, P, grid, nz);
        case 7: return launch_tma<7, 2>(e, P, grid, nz);
        case 8: return launch_tma<8, 2>(e, P, grid, nz);
        }
#undef S3D_TMA_CASE
    }
    const size_t smem = smem_bytes(hw, nz);
    switch (hw) {
    case 1: return launch<1>(e, P, grid, smem);
    case 2: return launch<2>(e, P, grid, smem);
    case 3: return launch<3>(e, P, grid, smem);
    case 4: return launch<4>(e, P, grid, smem);
    case 5: return launch<5>(e, P, grid, smem);
    case 6: return launch<6>(e, P, grid, smem);
    case 7: return launch<7>(e, P, grid, smem);
    case 8: return launch<8>(e, P, grid, smem);
    }
    return s3d_fail(e, "fused blur: unsupported width", cudaSuccess, __FILE__, __LINE__);
}
