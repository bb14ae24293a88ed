// desc_cell_geom.cuh -- window geometry of the cell-owner descriptor kernel (k_descriptor3).
//
// extract_descrip (sift.c:1834-1928) visits every voxel of the window sphere whose rotated,
// scaled coordinate vb = (R^T v + half) * bin_fctr lies in [0, 4)^3 and scatters it into the
// 2x2x2 spatial cells around vb (SIFT3D_desc_acc_interp, sift.c:1687-1791).  k_descriptor3
// enumerates those voxels BY BASE CELL: lane l of a warp owns the base cell ib = floor(vb) =
// (l & 3, (l >> 2) & 3, 2 * pass + (l >> 4)) for the whole kernel, so the 32 lanes of a warp
// always scatter into 32 different cells -- with the vertex-major histogram layout of the kernel
// that is 32 different shared-memory banks, i.e. conflict-free atomics.
//
// A cell is a cube of side hist_width in the rotated frame.  The functions below give
//   * d3_cell_bbox: a (y, z) bounding box of the cell in voxel rows (a superset), and
//   * d3_scan_row:  for one row, an x interval that CONTAINS every voxel of the row that
//                   belongs to the cell and the sphere (approximate arithmetic, widened),
//   * d3_member:    the exact per-voxel test and coordinates, in the reference's f32
//                   operation order (separately rounded multiplies and adds).
// Every voxel of a scanned interval is passed through d3_member, so the intervals only have to
// be supersets; every voxel the reference visits has exactly one base cell, so it is visited
// exactly once.  The code is __host__ __device__: tools/desc_cell_host_check.cu runs it on the
// CPU against a brute-force sweep of the window (tests/test_desc_cell_host.py).
#pragma once

#include <cuda_runtime.h>

#include <cmath>

#ifdef __CUDA_ARCH__
#define D3_MUL(a, b) __fmul_rn((a), (b))
#define D3_ADD(a, b) __fadd_rn((a), (b))
#define D3_SUB(a, b) __fsub_rn((a), (b))
#else
static inline float d3h_mul(float a, float b) { volatile float r = a * b; return r; }
static inline float d3h_add(float a, float b) { volatile float r = a + b; return r; }
static inline float d3h_sub(float a, float b) { volatile float r = a - b; return r; }
#define D3_MUL(a, b) d3h_mul((a), (b))
#define D3_ADD(a, b) d3h_add((a), (b))
#define D3_SUB(a, b) d3h_sub((a), (b))
#endif

// Per-keypoint constants (block-uniform).
struct D3Key {
    float kx, ky, kz;     // keypoint centre, voxels of the (local) level buffer
    float ux, uy, uz;     // (float) units of the level
    float Rt[9];          // R^T, row-major (sift.c:1854-1857)
    float half, binf;     // half descriptor width, 1 / hist_width (sift.c:1845-1850)
    float r2;             // win_radius^2
    float hw;             // hist_width
    int x0, x1, y0, y1, z0, z1;  // IM_LOOP_SPHERE bounds, clamped to [1, n-2] (sift.c:96-119)
    // derived by d3_key_finish():
    float isl[3];         // 1 / (d vb_a / d x), 0 where the slab does not depend on x
    int flat;             // bit a: |Rt[3a]| < 0.01 (vb_a varies by < 0.03 along any row)
    int ortho;            // R^T R = I to 1e-3 (the cell bounding boxes below assume it)
};

struct D3Cell {
    float ibf[3];            // (float) base cell index per rotated axis
    int ylo, yhi, zlo, zhi;  // rows that can hold voxels of the cell (may be empty: lo > hi)
};

__host__ __device__ inline void d3_key_finish(D3Key &K)
{
    K.flat = 0;
    for (int a = 0; a < 3; a++) {
        const float r = K.Rt[3 * a];
        if (fabsf(r) < 0.01f) {
            K.flat |= 1 << a;
            K.isl[a] = 0.0f;
        } else {
            K.isl[a] = 1.0f / (r * K.ux * K.binf);
        }
    }
    float dev = 0.0f;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            float s = 0.0f;
            for (int a = 0; a < 3; a++) s += K.Rt[3 * a + i] * K.Rt[3 * a + j];
            dev = fmaxf(dev, fabsf(s - (i == j ? 1.0f : 0.0f)));
        }
    K.ortho = dev < 1e-3f;
}

// Bounding rows of base cell (i0, i1, i2): q_a in [i_a * hw - half, (i_a + 1) * hw - half) in
// the rotated frame; v = Rt^T q for an orthonormal Rt, so along image axis i the cell spans
// c_i +- (hw / 2) * sum_a |Rt[3a + i]|.  Widened by 0.1 voxel (R^T R = I holds to
// 1e-3 and the window radius is below 100 voxels); the whole window if R is not
// orthonormal (user-supplied keypoints: still correct, only slower).
__host__ __device__ inline void d3_cell_bbox(const D3Key &K, int i0, int i1, int i2, D3Cell &C)
{
    C.ibf[0] = (float)i0, C.ibf[1] = (float)i1, C.ibf[2] = (float)i2;
    if (!K.ortho) {
        C.ylo = K.y0, C.yhi = K.y1, C.zlo = K.z0, C.zhi = K.z1;
        return;
    }
    const float q0 = ((float)i0 + 0.5f) * K.hw - K.half, q1 = ((float)i1 + 0.5f) * K.hw - K.half,
                q2 = ((float)i2 + 0.5f) * K.hw - K.half;
    const float cy = K.Rt[1] * q0 + K.Rt[4] * q1 + K.Rt[7] * q2;
    const float cz = K.Rt[2] * q0 + K.Rt[5] * q1 + K.Rt[8] * q2;
    const float ey = 0.5f * K.hw * (fabsf(K.Rt[1]) + fabsf(K.Rt[4]) + fabsf(K.Rt[7]));
    const float ez = 0.5f * K.hw * (fabsf(K.Rt[2]) + fabsf(K.Rt[5]) + fabsf(K.Rt[8]));
    const int ylo = (int)ceilf(K.ky + (cy - ey) / K.uy - 0.1f), yhi = (int)floorf(K.ky + (cy + ey) / K.uy + 0.1f);
    const int zlo = (int)ceilf(K.kz + (cz - ez) / K.uz - 0.1f), zhi = (int)floorf(K.kz + (cz + ez) / K.uz + 0.1f);
    C.ylo = ylo > K.y0 ? ylo : K.y0;
    C.yhi = yhi < K.y1 ? yhi : K.y1;
    C.zlo = zlo > K.z0 ? zlo : K.z0;
    C.zhi = zhi < K.z1 ? zhi : K.z1;
}

// Row-scan constants of a keypoint (block-uniform; the kernel keeps them in shared memory).
struct D3Scan {
    float kx, ky, kz, uy, uz, iux;
    float r2w;             // r2 * (1 + 2e-5)
    float b1[3], b2[3];    // Rt[3a + 1] * binf, Rt[3a + 2] * binf
    float hb;              // half * binf
    float isl[3];
    int flat;
    int x0, x1;
};

__host__ __device__ inline void d3_scan_setup(const D3Key &K, D3Scan &Q)
{
    Q.kx = K.kx, Q.ky = K.ky, Q.kz = K.kz, Q.uy = K.uy, Q.uz = K.uz, Q.iux = 1.0f / K.ux;
    Q.r2w = K.r2 + 2e-5f * K.r2;
    for (int a = 0; a < 3; a++) {
        Q.b1[a] = K.Rt[3 * a + 1] * K.binf;
        Q.b2[a] = K.Rt[3 * a + 2] * K.binf;
        Q.isl[a] = K.isl[a];
    }
    Q.hb = K.half * K.binf;
    Q.flat = K.flat;
    Q.x0 = K.x0, Q.x1 = K.x1;
}

// x interval [xa, xa + cnt) of row (y, z) containing every voxel of the cell inside the sphere.
// Approximate arithmetic: the sphere chord is widened by 2e-5 * r2 under the root (its rounding
// error is about 4 ulp of r2 and the root magnifies it near tangent rows), the slab ends by
// 0.01 voxel (their error is below 3e-3 voxel for |Rt[3a]| >= 0.01: an ulp of vb divided by the
// slope), and a slab with |Rt[3a]| < 0.01 (vb_a then varies by less than 2.83 * 0.01 bins along a
// row: bin_fctr * win_radius = 2 sqrt 2) only rejects rows that miss it by more than 0.04 bin.
__host__ __device__ __forceinline__ void d3_scan_row(const D3Scan &Q, const float ibf[3], int y, int z,
                                                     int &xa, int &cnt)
{
    cnt = 0;
    xa = 0;
    const float vy = ((float)y - Q.ky) * Q.uy, vz = ((float)z - Q.kz) * Q.uz;
    const float rem = Q.r2w - fmaf(vy, vy, vz * vz);
    if (rem < 0.0f) return;
#ifdef __CUDA_ARCH__
    float rs;  // MUFU.RSQ, 2 ulp: far inside the widening (rem = 0 gives inf * 0: guarded)
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(fmaxf(rem, 1e-12f)));
    const float hx = rem * rs * Q.iux;
#else
    const float hx = sqrtf(rem) * Q.iux;
#endif
    float lo = -hx, hi = hx;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const float off = fmaf(Q.b1[a], vy, fmaf(Q.b2[a], vz, Q.hb)) - ibf[a];
        if (Q.flat & (1 << a)) {
            if (off < -0.04f || off > 1.04f) return;
        } else {
            const float t0 = -off * Q.isl[a], t1 = t0 + Q.isl[a];
            lo = fmaxf(lo, fminf(t0, t1));
            hi = fminf(hi, fmaxf(t0, t1));
        }
    }
    if (!(lo <= hi + 0.02f)) return;
    int a0 = (int)ceilf(Q.kx + lo - 0.01f), b0 = (int)floorf(Q.kx + hi + 0.01f);
    if (a0 < Q.x0) a0 = Q.x0;
    if (b0 > Q.x1) b0 = Q.x1;
    if (b0 < a0) return;
    xa = a0;
    cnt = b0 - a0 + 1;
}

// The exact per-voxel geometry (sift.c:1866-1881, 1700-1716): squared distance, the three bin
// coordinates, and whether the reference visits the voxel AND its base cell is ibf.  dv = vb - ib
// is exact for vb in [ib, ib + 1) (Sterbenz), so `dv >= 0 && dv < 1` is floor(vb) == ib.
// (r2 is passed separately: the kernel reads it back from shared memory, see k_descriptor2.)
__host__ __device__ __forceinline__ bool d3_member(const D3Key &K, const float ibf[3], float r2, float xf,
                                                   float yf, float zf, float &sq, float dv[3])
{
    const float vx = D3_MUL(D3_SUB(xf, K.kx), K.ux);
    const float vy = D3_MUL(D3_SUB(yf, K.ky), K.uy);
    const float vz = D3_MUL(D3_SUB(zf, K.kz), K.uz);
    sq = D3_ADD(D3_ADD(D3_MUL(vx, vx), D3_MUL(vy, vy)), D3_MUL(vz, vz));
    bool ok = !(sq > r2);
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const float vk = D3_ADD(D3_ADD(D3_MUL(K.Rt[3 * a], vx), D3_MUL(K.Rt[3 * a + 1], vy)),
                                D3_MUL(K.Rt[3 * a + 2], vz));
        const float vb = D3_MUL(D3_ADD(vk, K.half), K.binf);
        dv[a] = D3_SUB(vb, ibf[a]);
        ok = ok && dv[a] >= 0.0f && dv[a] < 1.0f;
    }
    return ok;
}
