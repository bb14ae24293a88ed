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
    // ---- work decomposition: columns (overlapping last tile) x balanced z ranges,
    //      computed once per volume size and cached in the engine -------------------------
    const SegTab *tab = nullptr;
    for (const auto &t : e->segtabs)
        if (t.nx == nx && t.ny == ny && t.nz == nz && t.zb == zb && t.ze == ze) tab = &t;
    if (!tab) {
        std::vector<int> xs, ys;
        for (int x = 0; x < nx; x += TX) xs.push_back(std::min(x, nx - TX));
        for (int y = 0; y < ny; y += TY) ys.push_back(std::min(y, ny - TY));
        const long ncol = (long)xs.size() * ys.size();
        const long total = ncol * nzr;
        const int grid = (int)std::min<long>(e->num_sms, std::max<long>(1, total / 16));
        // Edge columns cost more per plane (mirror samples are synthesised by a few threads on the
        // critical path of every step); measured on B200 (tools/blur_dbg.py, cycles/step relative
        // to an interior column): left 1.11, right 1.30, top 1.06, bottom 1.19.  The z ranges are
        // balanced by cost so that all persistent CTAs finish together.
        std::vector<double> wcol(ncol);
        double wsum = 0;
        for (long c = 0; c < ncol; c++) {
            const int x0 = xs[c % xs.size()], y0 = ys[c / xs.size()];
            double w = 1.0;
            if (x0 - hw < 0) w *= e->blur_w[0];
            if (x0 + TX + hw > nx - 1) w *= e->blur_w[1];
            if (y0 - hw < 0) w *= e->blur_w[2];
            if (y0 + TY + hw > ny - 1) w *= e->blur_w[3];
            wcol[c] = w;
            wsum += w * nzr;
        }
        std::vector<Seg> segs;
        std::vector<int> start(grid + 1, 0);
        {
            const double quota = wsum / grid;
            long col = 0;
            int z = zb;
            for (int b = 0; b < grid; b++) {
                start[b] = (int)segs.size();
                double need = quota;
                while (col < ncol && (need > 1e-9 || b == grid - 1)) {
                    const double per = wcol[col];  // halo planes ignored
                    int take = (b == grid - 1) ? ze - z : (int)std::min<double>(ze - z, std::ceil(need / per - 1e-9));
                    if (take <= 0) break;
                    // avoid leaving a sliver shorter than the halo at the end of a column
                    if (ze - (z + take) > 0 && ze - (z + take) < 2 * hw && b != grid - 1) take = ze - z;
                    Seg sg;
                    sg.x0 = xs[col % xs.size()];
                    sg.y0 = ys[col / xs.size()];
                    sg.za = z;
                    sg.zb = z + take;
                    segs.push_back(sg);
                    need -= take * per;
                    z += take;
                    if (z >= ze) {
                        z = zb;
                        col++;
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
        nt.zb = zb, nt.ze = ze;
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
