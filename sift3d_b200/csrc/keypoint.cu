// keypoint.cu -- orientation assignment, sparse descriptors and dense descriptors.
//
// Replaces (reference bbrister/SIFT3D v1.4.6):
//   assign_orientations / assign_eig_ori   sift.c:1264-1514  (+ eigen_Mat_rm imutil.c:2992)
//   extract_descrip / SIFT3D_desc_acc_interp / icos_hist_bin / cart2bary / normalize_desc
//                                           sift.c:1834-1928, 1687-1791, 1646-1683, 335-394, 1794-1821
//   extract_dense_descriptors_no_rotate / postproc_Hist   sift.c:2429-2496, 2246-2292
//
// Precision rule: all geometry is evaluated in f32 in the
// reference's operation order with separately rounded multiplies and adds; only
// the ORDER of histogram accumulation differs (parallel), which costs ~1e-7
// relative L2 against a 1e-4 budget.
#include "common.cuh"
#include "desc_cell_geom.cuh"

#include <cfloat>
#include <cmath>

static inline float __fdiv_rn_host(float u) { volatile float r = 1.0f / u; return r; }

namespace {

__device__ __forceinline__ float fm(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fa(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fs(float a, float b) { return __fsub_rn(a, b); }
// a*b + c*d + e*f in the reference's left-to-right order
__device__ __forceinline__ float dot3(float a, float b, float c, float d, float e, float f)
{
    return fa(fa(fm(a, b), fm(c, d)), fm(e, f));
}

#define K_BARY_EPS ((double)FLT_EPSILON * 1E1)

// unit centroid directions of the 20 faces: uniform (unrolled) access -> constant bank operands
__constant__ float c_vmid[20][3];
// face index by (dodecahedron vertex family, sign bits of g): the 20 face centroids are the
// vertices (+-1,+-1,+-1), (+-1/phi,0,+-phi), (+-phi,+-1/phi,0), (0,+-phi,+-1/phi) of a dodecahedron
__constant__ int c_face_lut[32];

// Where the per-face constants and the preselection table live.  FaceMem: plain pointers (any
// kernel).  FaceSh: byte addresses in the shared window -- k_descriptor2 keeps ONE opaque base
// register for all of its shared data, because nvcc otherwise rematerialises the window base
// (S2R SR_CgaCtaId + MOV + LEA) at every access site under the register cap.
// Word k of a FaceConst: e1 0-2, e2 3-5, t 6-8, q 9-11, e2q 12, idx 13-15.
struct FaceMem {
    const FaceConst *F;
    const int *L;
    template <int K>
    __device__ __forceinline__ float w(int face) const
    {
        return reinterpret_cast<const float *>(F + face)[K];
    }
    template <int K>
    __device__ __forceinline__ int idx(int face) const { return F[face].idx[K]; }
    __device__ __forceinline__ int lut(int i) const { return L[i]; }
};
struct FaceSh {
    unsigned fa, la;  // shared-window byte addresses of FaceConst[20] and of the lut
    template <int K>
    __device__ __forceinline__ float w(int face) const
    {
        float v;
        asm("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(fa + (unsigned)face * (unsigned)sizeof(FaceConst)), "n"(4 * K));
        return v;
    }
    template <int K>
    __device__ __forceinline__ int idx(int face) const
    {
        int v;
        asm("ld.shared.s32 %0, [%1+%2];" : "=r"(v) : "r"(fa + (unsigned)face * (unsigned)sizeof(FaceConst)), "n"(4 * (13 + K)));
        return v;
    }
    __device__ __forceinline__ int lut(int i) const
    {
        int v;
        asm("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(la + 4u * (unsigned)i));
        return v;
    }
    // word at a run-time byte offset of a FaceConst (vertex indices in a rotated order)
    __device__ __forceinline__ int word(int face, unsigned byte_off) const
    {
        int v;
        asm("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(fa + (unsigned)face * (unsigned)sizeof(FaceConst) + byte_off));
        return v;
    }
};

// cart2bary (sift.c:335-394) with the per-face constants hoisted.
template <class FA>
__device__ __forceinline__ bool face_test(const FA &F, int f, const float g[3], float bary[3])
{
    const float e20 = F.template w<3>(f), e21 = F.template w<4>(f), e22 = F.template w<5>(f);
    float p[3];
    p[0] = fs(fm(g[1], e22), fm(g[2], e21));
    p[1] = fs(fm(g[2], e20), fm(g[0], e22));
    p[2] = fs(fm(g[0], e21), fm(g[1], e20));
    const float det = dot3(F.template w<0>(f), p[0], F.template w<1>(f), p[1], F.template w<2>(f), p[2]);
    if ((double)fabsf(det) < K_BARY_EPS) return false;
    const float det_inv = __fdiv_rn(1.0f, det);
    bary[1] = fm(det_inv, dot3(F.template w<6>(f), p[0], F.template w<7>(f), p[1], F.template w<8>(f), p[2]));
    bary[2] = fm(det_inv, dot3(g[0], F.template w<9>(f), g[1], F.template w<10>(f), g[2], F.template w<11>(f)));
    bary[0] = fs(fs(1.0f, bary[1]), bary[2]);
    const float k = fm(F.template w<12>(f), det_inv);
    return !((double)bary[0] < -K_BARY_EPS || (double)bary[1] < -K_BARY_EPS ||
             (double)bary[2] < -K_BARY_EPS || k < 0.0f);
}

// icos_hist_bin (sift.c:1646-1683): FIRST face in table order that passes.
// Fast path: the face whose three vertices have the largest dot products with g
// is the containing face; if its barycentrics are all comfortably positive no
// other face can pass (faces only overlap within bary_eps of shared edges), so
// it is also the first.  Otherwise fall back to the literal in-order loop.
template <class FA>
__device__ __forceinline__ int icos_bin_t(const FA &F, const float g[3], float bary[3],
                                          const bool fast = true)
{
    const float n2 = fa(fa(fm(g[0], g[0]), fm(g[1], g[1])), fm(g[2], g[2]));
    if ((double)n2 < K_BARY_EPS) return -1;
    if (fast) {
        // Preselect the face whose centroid is closest to g.  By symmetry only four of the twenty
        // centroid dot products can be the largest -- one per vertex family of the dual
        // dodecahedron, with the signs of g -- so 4 candidates replace 20 (the choice is then
        // VERIFIED with the exact test, so a wrong guess only costs the fallback loop).
        const float ax = fabsf(g[0]), ay = fabsf(g[1]), az = fabsf(g[2]);
        const float PHI = 1.6180339887f, IPH = 0.6180339887f;
        const float s0 = ax + ay + az, s1 = ax * IPH + az * PHI, s2 = ax * PHI + ay * IPH,
                    s3 = ay * PHI + az * IPH;
        int type = 0;
        float bs = s0;
        if (s1 > bs) bs = s1, type = 1;
        if (s2 > bs) bs = s2, type = 2;
        if (s3 > bs) bs = s3, type = 3;
        const int sb = (g[0] < 0.0f ? 1 : 0) | (g[1] < 0.0f ? 2 : 0) | (g[2] < 0.0f ? 4 : 0);
        const int best = F.lut(type * 8 + sb);
        const float margin = 1e-4f;
        if (face_test(F, best, g, bary) && bary[0] > margin && bary[1] > margin &&
            bary[2] > margin)
            return best;
    }
    for (int i = 0; i < 20; i++)
        if (face_test(F, i, g, bary)) return i;
    return -1;
}

// Face constants padded to 20 words (80 bytes) in shared memory: a face is fetched with four
// 16-byte loads instead of sixteen 4-byte ones.  Word k: e1 0-2, e2 3-5, t 6-8, q 9-11, e2q 12,
// vertex byte offsets 13-15 (k_descriptor3 rewrites the indices).
struct FaceSh4 {
    unsigned fa, la;  // shared-window byte addresses of the padded face array and of the lut
    __device__ __forceinline__ void load(int face, float4 v[4]) const
    {
        const unsigned a = fa + 80u * (unsigned)face;
        asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
            : "=f"(v[0].x), "=f"(v[0].y), "=f"(v[0].z), "=f"(v[0].w) : "r"(a));
        asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+16];"
            : "=f"(v[1].x), "=f"(v[1].y), "=f"(v[1].z), "=f"(v[1].w) : "r"(a));
        asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+32];"
            : "=f"(v[2].x), "=f"(v[2].y), "=f"(v[2].z), "=f"(v[2].w) : "r"(a));
        asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+48];"
            : "=f"(v[3].x), "=f"(v[3].y), "=f"(v[3].z), "=f"(v[3].w) : "r"(a));
    }
    __device__ __forceinline__ int lut(int i) const
    {
        int v;
        asm("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(la + 4u * (unsigned)i));
        return v;
    }
};

// face_test on a face already in registers (same arithmetic, same order)
__device__ __forceinline__ bool face_test4(const float4 v[4], const float g[3], float bary[3])
{
    const float e10 = v[0].x, e11 = v[0].y, e12 = v[0].z, e20 = v[0].w, e21 = v[1].x, e22 = v[1].y;
    const float t0 = v[1].z, t1 = v[1].w, t2 = v[2].x, q0 = v[2].y, q1 = v[2].z, q2 = v[2].w;
    float p[3];
    p[0] = fs(fm(g[1], e22), fm(g[2], e21));
    p[1] = fs(fm(g[2], e20), fm(g[0], e22));
    p[2] = fs(fm(g[0], e21), fm(g[1], e20));
    const float det = dot3(e10, p[0], e11, p[1], e12, p[2]);
    if ((double)fabsf(det) < K_BARY_EPS) return false;
    const float det_inv = __fdiv_rn(1.0f, det);
    bary[1] = fm(det_inv, dot3(t0, p[0], t1, p[1], t2, p[2]));
    bary[2] = fm(det_inv, dot3(g[0], q0, g[1], q1, g[2], q2));
    bary[0] = fs(fs(1.0f, bary[1]), bary[2]);
    const float k = fm(v[3].x, det_inv);
    return !((double)bary[0] < -K_BARY_EPS || (double)bary[1] < -K_BARY_EPS ||
             (double)bary[2] < -K_BARY_EPS || k < 0.0f);
}

// icos_bin_t on the padded table; also returns words 13-15 of the chosen face
template <bool FAST>
__device__ __forceinline__ int icos_bin4(const FaceSh4 &F, const float g[3], float bary[3], int voff[3])
{
    const float n2 = fa(fa(fm(g[0], g[0]), fm(g[1], g[1])), fm(g[2], g[2]));
    if ((double)n2 < K_BARY_EPS) return -1;
    float4 v[4];
    if (FAST) {
        const float ax = fabsf(g[0]), ay = fabsf(g[1]), az = fabsf(g[2]);
        const float PHI = 1.6180339887f, IPH = 0.6180339887f;
        const float s0 = ax + ay + az, s1 = ax * IPH + az * PHI, s2 = ax * PHI + ay * IPH,
                    s3 = ay * PHI + az * IPH;
        int type = 0;
        float bs = s0;
        if (s1 > bs) bs = s1, type = 1;
        if (s2 > bs) bs = s2, type = 2;
        if (s3 > bs) bs = s3, type = 3;
        const int sb = (g[0] < 0.0f ? 1 : 0) | (g[1] < 0.0f ? 2 : 0) | (g[2] < 0.0f ? 4 : 0);
        const int best = F.lut(type * 8 + sb);
        const float margin = 1e-4f;
        F.load(best, v);
        if (face_test4(v, g, bary) && bary[0] > margin && bary[1] > margin && bary[2] > margin) {
            voff[0] = __float_as_int(v[3].y), voff[1] = __float_as_int(v[3].z), voff[2] = __float_as_int(v[3].w);
            return best;
        }
    }
    for (int i = 0; i < 20; i++) {
        F.load(i, v);
        if (face_test4(v, g, bary)) {
            voff[0] = __float_as_int(v[3].y), voff[1] = __float_as_int(v[3].z), voff[2] = __float_as_int(v[3].w);
            return i;
        }
    }
    return -1;
}

__device__ __forceinline__ int icos_bin(const FaceConst *F /* shared memory */, const int *lut,
                                        const float g[3], float bary[3], const bool fast = true)
{
    const FaceMem A{F, lut};
    return icos_bin_t(A, g, bary, fast);
}

// stage the per-face constants in shared memory (divergent indexing by face)
__device__ __forceinline__ void load_faces(FaceConst *s_face, int *s_lut,
                                           const MeshDev *__restrict__ M)
{
    if (threadIdx.x < 32) s_lut[threadIdx.x] = c_face_lut[threadIdx.x];
    const int nw = 20 * (int)(sizeof(FaceConst) / 4);
    const unsigned *src = reinterpret_cast<const unsigned *>(M->f);
    unsigned *dst = reinterpret_cast<unsigned *>(s_face);
    for (int i = threadIdx.x; i < nw; i += blockDim.x) dst[i] = __ldg(src + i);
}

// cyclic Jacobi, f64, ascending eigenvalues, eigenvectors in the columns of Q
// (stand-in for LAPACK dsyevd, imutil.c:2992-3075; same algorithm as the oracle
// port so the two agree bit for bit)
__device__ void eig3(const double Ain[9], double Q[9], double L[3])
{
    double A[3][3], V[3][3];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            A[i][j] = Ain[3 * i + j];
            V[i][j] = i == j ? 1.0 : 0.0;
        }
    for (int sweep = 0; sweep < 64; sweep++) {
        const double off = __dadd_rn(__dadd_rn(fabs(A[0][1]), fabs(A[0][2])), fabs(A[1][2]));
        if (off == 0.0) break;
#pragma unroll
        for (int p = 0; p < 2; p++)
#pragma unroll
            for (int q = p + 1; q < 3; q++) {
                if (A[p][q] == 0.0) continue;
                const double theta =
                    __ddiv_rn(__dsub_rn(A[q][q], A[p][p]), __dmul_rn(2.0, A[p][q]));
                const double t = __ddiv_rn(
                    theta >= 0 ? 1.0 : -1.0,
                    __dadd_rn(fabs(theta), __dsqrt_rn(__dadd_rn(__dmul_rn(theta, theta), 1.0))));
                const double cs = __ddiv_rn(1.0, __dsqrt_rn(__dadd_rn(__dmul_rn(t, t), 1.0)));
                const double sn = __dmul_rn(t, cs);
                const double app = A[p][p], aqq = A[q][q], apq = A[p][q];
                A[p][p] = __dsub_rn(app, __dmul_rn(t, apq));
                A[q][q] = __dadd_rn(aqq, __dmul_rn(t, apq));
                A[p][q] = A[q][p] = 0.0;
                const int r = 3 - p - q;
                {
                    const double arp = A[r][p], arq = A[r][q];
                    A[r][p] = A[p][r] = __dsub_rn(__dmul_rn(cs, arp), __dmul_rn(sn, arq));
                    A[r][q] = A[q][r] = __dadd_rn(__dmul_rn(sn, arp), __dmul_rn(cs, arq));
                }
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const double vrp = V[k][p], vrq = V[k][q];
                    V[k][p] = __dsub_rn(__dmul_rn(cs, vrp), __dmul_rn(sn, vrq));
                    V[k][q] = __dadd_rn(__dmul_rn(sn, vrp), __dmul_rn(cs, vrq));
                }
            }
    }
    int order[3] = {0, 1, 2};
    double dg[3] = {A[0][0], A[1][1], A[2][2]};
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = i + 1; j < 3; j++)
            if (dg[order[j]] < dg[order[i]]) {
                const int t = order[i];
                order[i] = order[j];
                order[j] = t;
            }
#pragma unroll
    for (int i = 0; i < 3; i++) {
        L[i] = dg[order[i]];
#pragma unroll
        for (int j = 0; j < 3; j++) Q[3 * j + i] = V[j][order[i]];
    }
}

struct PyrTable {
    float *const *ptrs;  // gpyr level pointers [o*nlev_g + s+1]
    const int *dims;     // 3 per level
    const float *units;  // 3 per level ((float) ux, uy, uz)
    const double *scales;
    int nlev_g;
    int first_level;
    const int *zoff;  // Z-slab tiling: global z of local plane 0, per level (nullptr: whole volumes)
    const float4 *const *gptrs;  // gradient volumes per level (nullptr table or entries: none)
};

// Keypoint z coordinates are global; the level buffers of a Z-slab engine start at plane
// zoff.  The difference of two integer-valued floats is exact, and every later use of the
// centre is relative (voxel - centre), so tiled and untiled runs see identical arithmetic.
__device__ __forceinline__ float local_z(const PyrTable &T, int lv, float z)
{
    return T.zoff ? __fsub_rn(z, (float)T.zoff[lv]) : z;
}

// IM_LOOP_SPHERE_START bounds (sift.c:96-119) with a double radius
__device__ __forceinline__ void sphere_bounds_d(float c, double rad, float uf, int n, int &lo,
                                                int &hi)
{
    lo = (int)fmaxf(floorf((float)((double)c - rad / (double)uf)), 1.0f);
    hi = (int)fminf(ceilf((float)((double)c + rad / (double)uf)), (float)(n - 2));
}
// ... and with a float radius
__device__ __forceinline__ void sphere_bounds_f(float c, float rad, float uf, int n, int &lo,
                                                int &hi)
{
    lo = (int)fmaxf(floorf(fs(c, __fdiv_rn(rad, uf))), 1.0f);
    hi = (int)fminf(ceilf(fa(c, __fdiv_rn(rad, uf))), (float)(n - 2));
}

// expf exactly as glibc >= 2.27 computes it (sysdeps/ieee754/flt-32/e_expf.c: N = 32 table,
// degree-3 polynomial in f64, result narrowed to f32), for |x| < 88.  Checked on the host
// against libm's expf: 0 mismatches in 2e7 random arguments in [-2.2, 0], for both the
// FMA and the non-FMA evaluation order (tools/expf_check.c).  A merely "correctly rounded"
// exp differs from glibc in 0.06 % of the calls, CUDA's 2-ulp expf in far more; a 1-ulp
// change of a window weight can flip the icosahedron face of a gradient that sits within
// bary_eps of an edge, so the weights are reproduced bit for bit instead.
__constant__ unsigned long long c_exp2f_tab[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull};

// TabPtr / TabSh: the 32-entry table through a pointer or a shared-window byte address
struct TabPtr {
    const unsigned long long *t;
    __device__ __forceinline__ unsigned long long operator()(unsigned i) const { return t[i]; }
};
struct TabSh {
    unsigned a;
    __device__ __forceinline__ unsigned long long operator()(unsigned i) const
    {
        unsigned long long v;
        asm("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a + 8u * i));
        return v;
    }
};

template <class TAB>
__device__ __forceinline__ float expf_glibc_t(float x, const TAB &tab)
{
    const double InvLn2N = 0x1.71547652b82fep+0 * 32, SHIFT = 0x1.8p+52;
    const double C0 = 0x1.c6af84b912394p-5 / 32 / 32 / 32, C1 = 0x1.ebfce50fac4f3p-3 / 32 / 32,
                 C2 = 0x1.62e42ff0c52d6p-1 / 32;
    const double z = __dmul_rn(InvLn2N, (double)x);
    double kd = __dadd_rn(z, SHIFT);
    const unsigned long long ki = (unsigned long long)__double_as_longlong(kd);
    kd = __dsub_rn(kd, SHIFT);
    const double r = __dsub_rn(z, kd);
    const unsigned long long t = tab((unsigned)ki & 31u) + (ki << 47);
    const double sc = __longlong_as_double((long long)t);
    const double zz = __fma_rn(C0, r, C1);
    const double r2 = __dmul_rn(r, r);
    double y = __fma_rn(C2, r, 1.0);
    y = __fma_rn(zz, r2, y);
    y = __dmul_rn(y, sc);
    return (float)y;
}

__device__ __forceinline__ float expf_glibc(float x, const unsigned long long *tab /* shared */)
{
    return expf_glibc_t(x, TabPtr{tab});
}

// Second half of assign_eig_ori (sift.c:1424-1497): eigenvectors of the structure tensor,
// eigenvalue-ratio and corner tests, sign-fixed rotation matrix.
__device__ __forceinline__ bool orient_finish(double a00, double a01, double a02, double a11,
                                              double a12, double a22, float wx, float wy, float wz,
                                              double corner_thresh, float R[9], double &conf)
{
    bool accept = true;
    conf = 0.0;
#pragma unroll
    for (int k = 0; k < 9; k++) R[k] = 0.0f;
    const float wn2 = fa(fa(fm(wx, wx), fm(wy, wy)), fm(wz, wz));
    if (wn2 < (float)1E-10) accept = false;  // ori_grad_thresh, sift.c:1426
    if (accept) {
        const double A[9] = {a00, a01, a02, a01, a11, a12, a02, a12, a22};
        double Q[9], L[3];
        eig3(A, Q, L);
        if (fabs(__ddiv_rn(L[0], L[1])) > 0.90 || fabs(__ddiv_rn(L[1], L[2])) > 0.90)
            accept = false;  // max_eig_ratio, sift.c:1440-1444
        if (accept) {
            double corner = DBL_MAX;
            float v[2][3];
#pragma unroll
            for (int k = 0; k < 2; k++) {  // sift.c:1448-1480
                const int ei = 2 - k;
                float vr0 = (float)Q[0 * 3 + ei], vr1 = (float)Q[1 * 3 + ei],
                      vr2 = (float)Q[2 * 3 + ei];
                const double d = (double)dot3(wx, vr0, wy, vr1, wz, vr2);
                const float nv = __fsqrt_rn(fa(fa(fm(vr0, vr0), fm(vr1, vr1)), fm(vr2, vr2)));
                const float nw = __fsqrt_rn(wn2);
                const double cos_ang = __ddiv_rn(d, (double)fm(nv, nw));
                corner = fmin(corner, fabs(cos_ang));
                const float sgn = d > 0.0 ? 1.0f : -1.0f;
                vr0 = fm(vr0, sgn);
                vr1 = fm(vr1, sgn);
                vr2 = fm(vr2, sgn);
                R[0 * 3 + k] = vr0;
                R[1 * 3 + k] = vr1;
                R[2 * 3 + k] = vr2;
                v[k][0] = vr0;
                v[k][1] = vr1;
                v[k][2] = vr2;
            }
            R[0 * 3 + 2] = fs(fm(v[0][1], v[1][2]), fm(v[0][2], v[1][1]));
            R[1 * 3 + 2] = fs(fm(v[0][2], v[1][0]), fm(v[0][0], v[1][2]));
            R[2 * 3 + 2] = fs(fm(v[0][0], v[1][1]), fm(v[0][1], v[1][0]));
            conf = corner;
            if (corner < corner_thresh) accept = false;  // sift.c:1340-1341
        }
    }
    return accept;
}

// ---------------------------------------------------------------- orientation
// assign_eig_ori + the corner threshold (sift.c:1336-1497) for one window centre, walked in
// the reference's raster order so the f32 window gradient and the f64 structure tensor see
// the same rounding sequence as the CPU (SURVEY.md "Orientation accept/reject flips").
// Returns accept; R is all-zero when the window was rejected before the eigenvectors.
__device__ __forceinline__ bool orient_core(const float *__restrict__ im, int nx, int ny, int nz,
                                            float uxf, float uyf, float uzf, float vcx, float vcy,
                                            float vcz, double sigma, double corner_thresh,
                                            const unsigned long long *s_tab, float R[9],
                                            double &conf)
{
    const double win_radius = sigma * 3.0;  // ori_rad_fctr, sift.c:1365
    const double r2 = win_radius * win_radius;
    const double s2 = sigma * sigma;
    const float iux = __fdiv_rn(1.0f, uxf), iuy = __fdiv_rn(1.0f, uyf), iuz = __fdiv_rn(1.0f, uzf);
    int x0, x1, y0, y1, z0, z1;
    sphere_bounds_d(vcx, win_radius, uxf, nx, x0, x1);
    sphere_bounds_d(vcy, win_radius, uyf, ny, y0, y1);
    sphere_bounds_d(vcz, win_radius, uzf, nz, z0, z1);
    const size_t ys = nx, zs = (size_t)nx * ny;

    double a00 = 0, a01 = 0, a02 = 0, a11 = 0, a12 = 0, a22 = 0;
    float wx = 0.0f, wy = 0.0f, wz = 0.0f;
    for (int z = z0; z <= z1; z++) {
        const float dz = fm(fs((float)z, vcz), uzf);
        const float dz2 = fm(dz, dz);
        for (int y = y0; y <= y1; y++) {
            const float dy = fm(fs((float)y, vcy), uyf);
            const float dy2 = fm(dy, dy);
            const float *row = im + (size_t)y * ys + (size_t)z * zs;
            for (int x = x0; x <= x1; x++) {
                const float dx = fm(fs((float)x, vcx), uxf);
                const float sq = fa(fa(fm(dx, dx), dy2), dz2);
                if ((double)sq > r2) continue;
                // weight = expf(-0.5 * sq_dist / (sigma * sigma)), sift.c:1401: f64
                // argument narrowed to f32, then glibc's expf
                const float arg = (float)__ddiv_rn(__dmul_rn(-0.5, (double)sq), s2);
                const float w = expf_glibc(arg, s_tab);
                const float *p = row + x;
                float gx = fm(0.5f, fs(__ldg(p + 1), __ldg(p - 1)));
                float gy = fm(0.5f, fs(__ldg(p + ys), __ldg(p - ys)));
                float gz = fm(0.5f, fs(__ldg(p + zs), __ldg(p - zs)));
                gx = fm(gx, iux);
                gy = fm(gy, iuy);
                gz = fm(gz, iuz);
                const double dw = (double)w, gxd = gx, gyd = gy, gzd = gz;
                a00 = __dadd_rn(a00, __dmul_rn(__dmul_rn(gxd, gxd), dw));
                a01 = __dadd_rn(a01, __dmul_rn(__dmul_rn(gxd, gyd), dw));
                a02 = __dadd_rn(a02, __dmul_rn(__dmul_rn(gxd, gzd), dw));
                a11 = __dadd_rn(a11, __dmul_rn(__dmul_rn(gyd, gyd), dw));
                a12 = __dadd_rn(a12, __dmul_rn(__dmul_rn(gyd, gzd), dw));
                a22 = __dadd_rn(a22, __dmul_rn(__dmul_rn(gzd, gzd), dw));
                wx = fa(wx, fm(gx, w));
                wy = fa(wy, fm(gy, w));
                wz = fa(wz, fm(gz, w));
            }
        }
    }
    return orient_finish(a00, a01, a02, a11, a12, a22, wx, wy, wz, corner_thresh, R, conf);
}

// ---- window-weight tables for integer-centred candidates -----------------------------------
// Every detector candidate of a level (o, s) sits on an integer voxel and uses the same sigma,
// so the sphere test and the window weight expf(-0.5 * d^2 / sigma^2) (sift.c:1393-1401) depend
// only on the integer offset (dx, dy, dz): they are tabulated once per level by k_orient_table
// with exactly the arithmetic of orient_core (-1 marks offsets outside the sphere), which takes
// the f64 division and the exp out of the per-candidate loop.
struct OriTab {
    int off;  // first entry in the table pool (and in the list pool), -1 if the level has none
    int rx, ry, rz;
    int list_n;  // entries of the level's in-sphere offset list (k_orient_list)
    int pad[3];
};

__global__ void __launch_bounds__(256)
    k_orient_table(const OriTab *__restrict__ tabs, PyrTable T, double sig_fctr,
                   float *__restrict__ pool)
{
    __shared__ unsigned long long s_tab[32];
    if (threadIdx.x < 32) s_tab[threadIdx.x] = c_exp2f_tab[threadIdx.x];
    __syncthreads();
    const int lv = blockIdx.y;
    const OriTab t = tabs[lv];
    if (t.off < 0) return;
    const int wx = 2 * t.rx + 1, wy = 2 * t.ry + 1, wz = 2 * t.rz + 1;
    const double sigma = sig_fctr * T.scales[lv];
    const double r2 = (sigma * 3.0) * (sigma * 3.0), s2 = sigma * sigma;
    const float uxf = T.units[3 * lv], uyf = T.units[3 * lv + 1], uzf = T.units[3 * lv + 2];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < wx * wy * wz;
         i += gridDim.x * blockDim.x) {
        const int ix = i % wx - t.rx, iy = (i / wx) % wy - t.ry, iz = i / (wx * wy) - t.rz;
        const float dx = fm((float)ix, uxf), dy = fm((float)iy, uyf), dz = fm((float)iz, uzf);
        const float sq = fa(fa(fm(dx, dx), fm(dy, dy)), fm(dz, dz));
        float w = -1.0f;
        if (!((double)sq > r2))
            w = expf_glibc((float)__ddiv_rn(__dmul_rn(-0.5, (double)sq), s2), s_tab);
        pool[t.off + i] = w;
    }
}

// orient_core for an integer centre with a weight table: same voxels, same order, same values.
// BATCH = 4: four voxels fetched ahead; BATCH = 8: batches end on the 128-byte lines of the
// gradient volume, so the loads of a batch are ONE L2 request however little of the line
// survives in L1 between batches (1024 threads x their own lines thrash it).
template <int BATCH>
__device__ __forceinline__ bool orient_core_tab(const float *__restrict__ im, int nx, int ny,
                                                int nz, float uxf, float uyf, float uzf, int cx,
                                                int cy, int cz, double sigma,
                                                const float *__restrict__ tab, const OriTab t,
                                                const float4 *__restrict__ gim,
                                                double corner_thresh, float R[9], double &conf)
{
    const double win_radius = sigma * 3.0;
    const float iux = __fdiv_rn(1.0f, uxf), iuy = __fdiv_rn(1.0f, uyf), iuz = __fdiv_rn(1.0f, uzf);
    int x0, x1, y0, y1, z0, z1;
    sphere_bounds_d((float)cx, win_radius, uxf, nx, x0, x1);
    sphere_bounds_d((float)cy, win_radius, uyf, ny, y0, y1);
    sphere_bounds_d((float)cz, win_radius, uzf, nz, z0, z1);
    // the table covers offsets up to +-r*; the loop bounds never exceed ceil(radius / unit)
    x0 = max(x0, cx - t.rx), x1 = min(x1, cx + t.rx);
    y0 = max(y0, cy - t.ry), y1 = min(y1, cy + t.ry);
    z0 = max(z0, cz - t.rz), z1 = min(z1, cz + t.rz);
    const size_t ys = nx, zs = (size_t)nx * ny;
    const int twx = 2 * t.rx + 1, twy = 2 * t.ry + 1;
    double a00 = 0, a01 = 0, a02 = 0, a11 = 0, a12 = 0, a22 = 0;
    float wx = 0.0f, wy = 0.0f, wz = 0.0f;
    for (int z = z0; z <= z1; z++)
        for (int y = y0; y <= y1; y++) {
            const size_t roff = (size_t)y * ys + (size_t)z * zs;
            const float *row = im + roff;
            const float *trow = tab + ((size_t)(z - cz + t.rz) * twy + (y - cy + t.ry)) * twx + (t.rx - cx);
            // A thread walks its window alone (the sums must see the reference's order), so
            // its speed is set by load latency: fetch four voxels' weights and gradients
            // before consuming them in order.
            auto acc = [&](float w, float gx, float gy, float gz) {
                const double dw = (double)w, gxd = gx, gyd = gy, gzd = gz;
                a00 = __dadd_rn(a00, __dmul_rn(__dmul_rn(gxd, gxd), dw));
                a01 = __dadd_rn(a01, __dmul_rn(__dmul_rn(gxd, gyd), dw));
                a02 = __dadd_rn(a02, __dmul_rn(__dmul_rn(gxd, gzd), dw));
                a11 = __dadd_rn(a11, __dmul_rn(__dmul_rn(gyd, gyd), dw));
                a12 = __dadd_rn(a12, __dmul_rn(__dmul_rn(gyd, gzd), dw));
                a22 = __dadd_rn(a22, __dmul_rn(__dmul_rn(gzd, gzd), dw));
                wx = fa(wx, fm(gx, w));
                wy = fa(wy, fm(gy, w));
                wz = fa(wz, fm(gz, w));
            };
            if (gim) {
                const float4 *grow = gim + roff;
                for (int x = x0; x <= x1;) {
                    // BATCH 8: up to the end of the current 128-byte line (8 float4)
                    const int nb = BATCH == 8 ? min(8 - (int)((roff + (size_t)x) & 7), x1 - x + 1)
                                              : min(4, x1 - x + 1);
                    float w[BATCH];
                    float4 g4[BATCH];
#pragma unroll
                    for (int k = 0; k < BATCH; k++) w[k] = k < nb ? __ldg(trow + x + k) : -1.0f;
                    // gradients only inside the sphere (48 % of the box is outside): the window
                    // traffic, 16 B per visited voxel from L2, is what bounds this kernel
#pragma unroll
                    for (int k = 0; k < BATCH; k++)
                        g4[k] = w[k] >= 0.0f ? __ldg(grow + x + k) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int k = 0; k < BATCH; k++)
                        if (w[k] >= 0.0f) acc(w[k], g4[k].x, g4[k].y, g4[k].z);
                    x += nb;
                }
            } else {
                for (int x = x0; x <= x1; x++) {
                    const float w = __ldg(trow + x);
                    if (w < 0.0f) continue;  // outside the sphere
                    const float *p = row + x;
                    acc(w, fm(fm(0.5f, fs(__ldg(p + 1), __ldg(p - 1))), iux),
                        fm(fm(0.5f, fs(__ldg(p + ys), __ldg(p - ys))), iuy),
                        fm(fm(0.5f, fs(__ldg(p + zs), __ldg(p - zs))), iuz));
                }
            }
        }
    return orient_finish(a00, a01, a02, a11, a12, a22, wx, wy, wz, corner_thresh, R, conf);
}

// One thread per candidate.  Candidates adjacent in scan order share (o, s) and integer
// centres, so the lanes of a warp walk identical offsets in lock step.
// TAB = true: detector candidates only (integer centres, sd == level scale, a table for every
// keypoint level -- the launcher guarantees it); a separate instantiation so that the literal
// path's f64 exp does not cost the fast one its occupancy.
template <bool TAB, int BATCH = 4>
__global__ void __launch_bounds__(128, TAB ? (BATCH == 8 ? 6 : 8) : 4)
    k_orient(s3d_keypoint *__restrict__ kps, int n, PyrTable T, double sig_fctr,
             double corner_thresh, unsigned char *__restrict__ ok, double *__restrict__ conf_out,
             const OriTab *__restrict__ tabs, const float *__restrict__ pool)
{
    __shared__ unsigned long long s_tab[32];
    if (!TAB) {
        if (threadIdx.x < 32) s_tab[threadIdx.x] = c_exp2f_tab[threadIdx.x];
        __syncthreads();
    }
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const s3d_keypoint c = kps[i];
    const int lv = c.o * T.nlev_g + (c.s - T.first_level);
    // detector: sigma = ori_sig_fctr * key->sd (sift.c:1281); raw API: sigma = key_base.sd
    // (sift.c:1579)
    float R[9];
    double conf;
    const float zl = local_z(T, lv, c.z);
    bool accept;
    if (TAB)
        accept = orient_core_tab<BATCH>(T.ptrs[lv], T.dims[3 * lv], T.dims[3 * lv + 1], T.dims[3 * lv + 2],
                                 T.units[3 * lv], T.units[3 * lv + 1], T.units[3 * lv + 2],
                                 (int)c.x, (int)c.y, (int)zl, sig_fctr * c.sd, pool + tabs[lv].off,
                                 tabs[lv], T.gptrs ? T.gptrs[lv] : nullptr, corner_thresh, R, conf);
    else
        accept = orient_core(T.ptrs[lv], T.dims[3 * lv], T.dims[3 * lv + 1], T.dims[3 * lv + 2],
                             T.units[3 * lv], T.units[3 * lv + 1], T.units[3 * lv + 2], c.x, c.y,
                             zl, sig_fctr * c.sd, corner_thresh, s_tab, R, conf);
#pragma unroll
    for (int k = 0; k < 9; k++) kps[i].R[k] = R[k];
    ok[i] = accept ? 1 : 0;
    if (conf_out) conf_out[i] = conf;
}

// ---- grouped orientation kernel (the default) -----------------------------------------------
// Eight lanes per candidate, four candidates per warp.  The in-sphere offsets of a level are a
// LIST in the reference's raster order (k_orient_list compacts the weight table), so a group
// walks 8 consecutive list entries per step: no sphere test, every lane busy, and the 8
// gradient fetches of a group are consecutive voxels of a row -- one 128-byte line instead of
// the 8 scattered sectors of the thread-per-candidate kernel, which is bound by that latency.
//   * The f32 window gradient (vd_win, sift.c:1416-1417) decides accept / reject through the
//     corner score, and a re-associated sum moves it by 1e-7 -- about one flipped candidate in
//     two 512^3 volumes.  It is therefore summed in the reference's exact order: the lanes
//     leave their terms g * w in shared memory and lanes 0..2 of the group add the eight terms
//     of "their" component one after the other (voxels outside the volume, which the
//     reference's clamped loop bounds skip, add an exact +0).
//   * The f64 structure tensor is order-insensitive at the 1e-16 level (its eigenvectors are
//     rounded to f32, and no decision sits within 1e-15 of its threshold): per-lane partial
//     sums, combined with a 3-step butterfly.
#define ORI_WARPS 4
// SCALAR: central differences from the level itself (7 x 4-byte loads, 4 bytes of footprint per
// voxel) instead of the float4 gradient volume (one 16-byte load, 16 bytes of footprint)
template <int ORI_G, bool SCALAR>
__global__ void __launch_bounds__(32 * ORI_WARPS)
    k_orient_group(s3d_keypoint *__restrict__ kps, int n, PyrTable T, double sig_fctr,
                   double corner_thresh, unsigned char *__restrict__ ok,
                   const OriTab *__restrict__ tabs, const int2 *__restrict__ lists)
{
    __shared__ __align__(16) float s_t[ORI_WARPS][32 / ORI_G][3][ORI_G];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane / ORI_G, gl = lane % ORI_G;
    const int i = (blockIdx.x * ORI_WARPS + warp) * (32 / ORI_G) + grp;
    const bool valid = i < n;
    s3d_keypoint c;
    c.o = 0, c.s = T.first_level, c.x = c.y = c.z = 0.0f, c.sd = 1.0;
    if (valid) c = kps[i];
    const int lv = c.o * T.nlev_g + (c.s - T.first_level);
    const OriTab t = tabs[lv];
    const int nl = valid ? t.list_n : 0;
    const int2 *__restrict__ list = lists + t.off;
    const int nx = T.dims[3 * lv], ny = T.dims[3 * lv + 1], nz = T.dims[3 * lv + 2];
    const float4 *__restrict__ gim = SCALAR ? nullptr : T.gptrs[lv];
    const float *__restrict__ im = T.ptrs[lv];
    const float iux = __fdiv_rn(1.0f, T.units[3 * lv]), iuy = __fdiv_rn(1.0f, T.units[3 * lv + 1]),
                iuz = __fdiv_rn(1.0f, T.units[3 * lv + 2]);
    const size_t ys = nx, zs = (size_t)nx * ny;
    const int cx = (int)c.x, cy = (int)c.y, cz = (int)local_z(T, lv, c.z);
    double a00 = 0, a01 = 0, a02 = 0, a11 = 0, a12 = 0, a22 = 0;
    float wsum = 0.0f;  // lanes 0..2 of the group: x, y, z component of the window gradient
    // ORI_U batches of 8 entries per trip: all their gradient fetches are issued before the
    // first is used (the kernel is bound by the latency of these gathers -- ncu: 11 stalled
    // warps on the long scoreboard per issue with one batch in flight)
    constexpr int ORI_U = 4;
    const int nit = __reduce_max_sync(0xffffffffu, (nl + ORI_U * ORI_G - 1) / (ORI_U * ORI_G));
    float *const st = &s_t[warp][grp][0][0];
    for (int it = 0; it < nit; it++) {
        float4 g[ORI_U];
        float w[ORI_U];
#pragma unroll
        for (int k = 0; k < ORI_U; k++) {
            const int e = (it * ORI_U + k) * ORI_G + gl;
            w[k] = -1.0f;  // no voxel
            g[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e < nl) {
                const int2 ent = __ldg(list + e);
                const int x = cx + (int)(signed char)(ent.x & 0xff), y = cy + (int)(signed char)((ent.x >> 8) & 0xff),
                          z = cz + (int)(signed char)((ent.x >> 16) & 0xff);
                // IM_LOOP_SPHERE_START clamps the loops to [1, n - 2] (sift.c:96-119)
                if ((unsigned)(x - 1) <= (unsigned)(nx - 3) && (unsigned)(y - 1) <= (unsigned)(ny - 3) &&
                    (unsigned)(z - 1) <= (unsigned)(nz - 3)) {
                    w[k] = __int_as_float(ent.y);
                    const size_t idx = ((size_t)z * ny + y) * nx + x;
                    if (SCALAR) {  // SIFT3D_IM_GET_GRAD (immacros.h:105) / units, as k_gradient
                        const float *p = im + idx;
                        g[k].x = fm(fm(0.5f, fs(__ldg(p + 1), __ldg(p - 1))), iux);
                        g[k].y = fm(fm(0.5f, fs(__ldg(p + ys), __ldg(p - ys))), iuy);
                        g[k].z = fm(fm(0.5f, fs(__ldg(p + zs), __ldg(p - zs))), iuz);
                    } else {
                        g[k] = __ldg(gim + idx);
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < ORI_U; k++) {
            float tx = 0.0f, ty = 0.0f, tz = 0.0f;
            if (w[k] >= 0.0f) {
                const double dw = (double)w[k], gxd = g[k].x, gyd = g[k].y, gzd = g[k].z;
                a00 = __dadd_rn(a00, __dmul_rn(__dmul_rn(gxd, gxd), dw));
                a01 = __dadd_rn(a01, __dmul_rn(__dmul_rn(gxd, gyd), dw));
                a02 = __dadd_rn(a02, __dmul_rn(__dmul_rn(gxd, gzd), dw));
                a11 = __dadd_rn(a11, __dmul_rn(__dmul_rn(gyd, gyd), dw));
                a12 = __dadd_rn(a12, __dmul_rn(__dmul_rn(gyd, gzd), dw));
                a22 = __dadd_rn(a22, __dmul_rn(__dmul_rn(gzd, gzd), dw));
                tx = fm(g[k].x, w[k]);
                ty = fm(g[k].y, w[k]);
                tz = fm(g[k].z, w[k]);
            }
            st[gl] = tx;
            st[ORI_G + gl] = ty;
            st[2 * ORI_G + gl] = tz;
            __syncwarp();
            if (gl < 3) {  // the reference's order: voxel after voxel (sift.c:1416-1417)
#pragma unroll
                for (int q = 0; q < ORI_G / 4; q++) {
                    const float4 v = *reinterpret_cast<const float4 *>(st + ORI_G * gl + 4 * q);
                    wsum = fa(wsum, v.x);
                    wsum = fa(wsum, v.y);
                    wsum = fa(wsum, v.z);
                    wsum = fa(wsum, v.w);
                }
            }
            __syncwarp();
        }
    }
#pragma unroll
    for (int o = 1; o < ORI_G; o <<= 1) {
        a00 = __dadd_rn(a00, __shfl_xor_sync(0xffffffffu, a00, o));
        a01 = __dadd_rn(a01, __shfl_xor_sync(0xffffffffu, a01, o));
        a02 = __dadd_rn(a02, __shfl_xor_sync(0xffffffffu, a02, o));
        a11 = __dadd_rn(a11, __shfl_xor_sync(0xffffffffu, a11, o));
        a12 = __dadd_rn(a12, __shfl_xor_sync(0xffffffffu, a12, o));
        a22 = __dadd_rn(a22, __shfl_xor_sync(0xffffffffu, a22, o));
    }
    const float wx = __shfl_sync(0xffffffffu, wsum, grp * ORI_G + 0);
    const float wy = __shfl_sync(0xffffffffu, wsum, grp * ORI_G + 1);
    const float wz = __shfl_sync(0xffffffffu, wsum, grp * ORI_G + 2);
    if (valid && gl == 0) {
        float R[9];
        double conf;
        const bool accept = orient_finish(a00, a01, a02, a11, a12, a22, wx, wy, wz, corner_thresh, R, conf);
#pragma unroll
        for (int k = 0; k < 9; k++) kps[i].R[k] = R[k];
        ok[i] = accept ? 1 : 0;
    }
}

// In-sphere offsets of every level's weight table, in raster order: {dx | dy << 8 | dz << 16
// (signed bytes), weight bits}.  One block per level; ordered compaction chunk by chunk.
__global__ void __launch_bounds__(1024)
    k_orient_list(OriTab *__restrict__ tabs, const float *__restrict__ pool, int2 *__restrict__ lists)
{
    __shared__ int s_warp[32];
    __shared__ int s_run;
    OriTab t = tabs[blockIdx.x];
    if (t.off < 0) return;
    const int wx = 2 * t.rx + 1, wy = 2 * t.ry + 1, wz = 2 * t.rz + 1;
    const int tot = wx * wy * wz;
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    for (int base = 0; base < tot; base += 1024) {
        const int i = base + threadIdx.x;
        const float w = i < tot ? pool[t.off + i] : -1.0f;
        const int v = w >= 0.0f ? 1 : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += u;
        }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            int ws = s_warp[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, ws, o);
                if (threadIdx.x >= o) ws += u;
            }
            s_warp[threadIdx.x] = ws;
        }
        __syncthreads();
        const int warp_off = (threadIdx.x >> 5) ? s_warp[(threadIdx.x >> 5) - 1] : 0;
        const int run = s_run;
        if (v) {
            const int ix = i % wx - t.rx, iy = (i / wx) % wy - t.ry, iz = i / (wx * wy) - t.rz;
            lists[t.off + run + warp_off + incl - 1] =
                make_int2((ix & 0xff) | ((iy & 0xff) << 8) | ((iz & 0xff) << 16), __float_as_int(w));
        }
        __syncthreads();
        if (threadIdx.x == 1023) s_run = run + warp_off + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) tabs[blockIdx.x].list_n = s_run;
}

// ordered compaction of accepted candidates (sift.c:1306-1324)
__global__ void __launch_bounds__(1024)
    k_flag_scan(const unsigned char *__restrict__ ok, int n, int *__restrict__ pos,
                int *__restrict__ counter)
{
    __shared__ int s_warp[32];
    __shared__ int s_run;
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < n ? ok[i] : 0;
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
        if (i < n) pos[i] = run + warp_off + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_run = run + warp_off + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) counter[1] = s_run;
}

__global__ void __launch_bounds__(256)
    k_cand_to_keypoints(const Candidate *__restrict__ cand, int n, const double *__restrict__ scales,
                        int nlev_g, int first_level, s3d_keypoint *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Candidate c = cand[i];
    s3d_keypoint k;
#pragma unroll
    for (int j = 0; j < 9; j++) k.R[j] = 0.0f;
    k.x = (float)c.x;  // key->xd = (double) x (sift.c:1203) -> Cvec float (sift.c:1280)
    k.y = (float)c.y;
    k.z = (float)c.z;
    k.sd = scales[c.o * nlev_g + (c.s - first_level)];  // key->sd = cur->s (sift.c:1202)
    k.o = c.o;
    k.s = c.s;
    out[i] = k;
}

__global__ void __launch_bounds__(256)
    k_compact_keypoints(const s3d_keypoint *__restrict__ in, const unsigned char *__restrict__ ok,
                        const int *__restrict__ pos, int n, s3d_keypoint *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !ok[i]) return;
    out[pos[i]] = in[i];
}

// ---------------------------------------------------------------- sparse descriptor
// One CTA per keypoint; the threads sweep the bounding box of the window sphere.
#define DESC_THREADS 256

__global__ void __launch_bounds__(DESC_THREADS)
    k_descriptor(const s3d_keypoint *__restrict__ kps, int n, PyrTable T,
                 const MeshDev *__restrict__ M, unsigned char *__restrict__ out, int icos_fast)
{
    __shared__ float hist[S3D_DESC_NUMEL];
    __shared__ double s_red[DESC_THREADS / 32];
    __shared__ float s_norm_inv;
    __shared__ unsigned long long s_tab[32];
    __shared__ FaceConst s_face[20];
    __shared__ int s_lut[32];
    const int ki = blockIdx.x;
    if (ki >= n) return;
    if (threadIdx.x < 32) s_tab[threadIdx.x] = c_exp2f_tab[threadIdx.x];
    load_faces(s_face, s_lut, M);
    s3d_keypoint kp = kps[ki];
    const int lv = kp.o * T.nlev_g + (kp.s - T.first_level);
    const float z_global = kp.z;
    kp.z = local_z(T, lv, kp.z);
    const float *__restrict__ im = T.ptrs[lv];
    const int nx = T.dims[3 * lv], ny = T.dims[3 * lv + 1], nz = T.dims[3 * lv + 2];
    const float uxf = T.units[3 * lv], uyf = T.units[3 * lv + 1], uzf = T.units[3 * lv + 2];
    const float iux = __fdiv_rn(1.0f, uxf), iuy = __fdiv_rn(1.0f, uyf), iuz = __fdiv_rn(1.0f, uzf);
    // sift.c:1845-1850
    const float sigma = (float)__dmul_rn(kp.sd, 7.071067812);
    const float win_radius = (float)__dmul_rn(2.0, (double)sigma);
    const float half = (float)__ddiv_rn((double)win_radius, sqrt(2.0));
    const float desc_width = fm(2.0f, half);
    const float hist_width = __fdiv_rn(desc_width, 4.0f);
    const float bin_fctr = __fdiv_rn(1.0f, hist_width);
    const float r2 = fm(win_radius, win_radius);
    const float s2 = fm(sigma, sigma);
    int x0, x1, y0, y1, z0, z1;
    sphere_bounds_f(kp.x, win_radius, uxf, nx, x0, x1);
    sphere_bounds_f(kp.y, win_radius, uyf, ny, y0, y1);
    sphere_bounds_f(kp.z, win_radius, uzf, nz, z0, z1);
    // Rt = R^T (sift.c:1854-1857)
    float Rt[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) Rt[3 * i + j] = kp.R[3 * j + i];

    for (int i = threadIdx.x; i < S3D_DESC_NUMEL; i += DESC_THREADS) hist[i] = 0.0f;
    __syncthreads();

    const int bx = max(x1 - x0 + 1, 0), by = max(y1 - y0 + 1, 0), bz = max(z1 - z0 + 1, 0);
    const int vol = bx * by * bz;
    const size_t ys = nx, zs = (size_t)nx * ny;
    for (int t = threadIdx.x; t < vol; t += DESC_THREADS) {
        const int x = x0 + t % bx;
        const int r = t / bx;
        const int y = y0 + r % by;
        const int z = z0 + r / by;
        const float vx = fm(fs((float)x, kp.x), uxf);
        const float vy = fm(fs((float)y, kp.y), uyf);
        const float vz = fm(fs((float)z, kp.z), uzf);
        const float sq = fa(fa(fm(vx, vx), fm(vy, vy)), fm(vz, vz));
        if (sq > r2) continue;
        float vb[3];
        bool inside = true;
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const float vk = dot3(Rt[3 * a], vx, Rt[3 * a + 1], vy, Rt[3 * a + 2], vz);
            vb[a] = fm(fa(vk, half), bin_fctr);
            inside = inside && !(vb[a] < 0.0f || vb[a] >= 4.0f);
        }
        if (!inside) continue;
        const float *p = im + x + (size_t)y * ys + (size_t)z * zs;
        float g[3];
        g[0] = fm(fm(0.5f, fs(__ldg(p + 1), __ldg(p - 1))), iux);
        g[1] = fm(fm(0.5f, fs(__ldg(p + ys), __ldg(p - ys))), iuy);
        g[2] = fm(fm(0.5f, fs(__ldg(p + zs), __ldg(p - zs))), iuz);
        // sift.c:1890: expf(-0.5f * sq_dist / (sigma * sigma)), f32 argument, glibc's expf
        const float w = expf_glibc(__fdiv_rn(fm(-0.5f, sq), s2), s_tab);
        g[0] = fm(g[0], w);
        g[1] = fm(g[1], w);
        g[2] = fm(g[2], w);
        float gr[3];
#pragma unroll
        for (int a = 0; a < 3; a++)
            gr[a] = dot3(Rt[3 * a], g[0], Rt[3 * a + 1], g[1], Rt[3 * a + 2], g[2]);
        float bary[3];
        const int bin = icos_bin(s_face, s_lut, gr, bary, icos_fast != 0);
        if (bin < 0) continue;
        const float mag = __fsqrt_rn(fa(fa(fm(gr[0], gr[0]), fm(gr[1], gr[1])), fm(gr[2], gr[2])));
        float dv[3];
        int ib[3];
#pragma unroll
        for (int a = 0; a < 3; a++) {
            dv[a] = fs(vb[a], floorf(vb[a]));
            ib[a] = (int)vb[a];
        }
        const int i0 = s_face[bin].idx[0], i1 = s_face[bin].idx[1], i2 = s_face[bin].idx[2];
#pragma unroll
        for (int dx = 0; dx < 2; dx++)
#pragma unroll
            for (int dy = 0; dy < 2; dy++)
#pragma unroll
                for (int dz = 0; dz < 2; dz++) {
                    const int cx = ib[0] + dx, cy = ib[1] + dy, cz = ib[2] + dz;
                    if (cx >= 4 || cy >= 4 || cz >= 4) continue;  // lower bounds hold (vb >= 0)
                    const float wgt = fm(fm(dx ? dv[0] : fs(1.0f, dv[0]), dy ? dv[1] : fs(1.0f, dv[1])),
                                         dz ? dv[2] : fs(1.0f, dv[2]));
                    float *h = hist + 12 * (cx + 4 * cy + 16 * cz);
                    const float mw = fm(mag, wgt);
                    atomicAdd(h + i0, fm(mw, bary[0]));
                    atomicAdd(h + i1, fm(mw, bary[1]));
                    atomicAdd(h + i2, fm(mw, bary[2]));
                }
    }
    __syncthreads();

    // normalize_desc (sift.c:1794-1821), truncate (sift.c:1909-1915), normalize again
    const float trunc = (float)((double)(0.2f * 128.0f / S3D_DESC_NUMEL));
    for (int pass = 0; pass < 2; pass++) {
        double acc = 0.0;
        for (int i = threadIdx.x; i < S3D_DESC_NUMEL; i += DESC_THREADS) {
            const double v = (double)hist[i];
            acc = __dadd_rn(acc, __dmul_rn(v, v));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double tot = 0.0;
            for (int i = 0; i < DESC_THREADS / 32; i++) tot += s_red[i];
            const double norm = sqrt(tot) + DBL_EPSILON;
            s_norm_inv = (float)(1.0 / norm);
        }
        __syncthreads();
        const float ninv = s_norm_inv;
        for (int i = threadIdx.x; i < S3D_DESC_NUMEL; i += DESC_THREADS) {
            float v = fm(hist[i], ninv);
            if (pass == 0) v = fminf(v, trunc);
            hist[i] = v;
        }
        __syncthreads();
    }
    float *o32 = reinterpret_cast<float *>(out + (size_t)ki * S3D_DESC_STRIDE);
    for (int i = threadIdx.x; i < S3D_DESC_NUMEL; i += DESC_THREADS) o32[i] = hist[i];
    if (threadIdx.x == 0) {
        double *o64 = reinterpret_cast<double *>(out + (size_t)ki * S3D_DESC_STRIDE +
                                                 S3D_DESC_NUMEL * sizeof(float));
        const double f = ldexp(1.0, kp.o);  // sift.c:1851, 1922-1925
        o64[0] = (double)kp.x * f;
        o64[1] = (double)kp.y * f;
        o64[2] = (double)z_global * f;
        o64[3] = kp.sd;
    }
}

// native 32-bit shared-memory atomics on a shared-window byte address
__device__ __forceinline__ unsigned atoms_add(unsigned addr, unsigned v)
{
    unsigned old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ void atoms_inc(unsigned addr)
{
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory");
}
__device__ __forceinline__ float ld_shared_f32(unsigned addr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
// mask = 2 * mask + carry_out(old + q): the carry flag travels inside ONE asm statement
__device__ __forceinline__ void carry_push(unsigned &mask, unsigned old, unsigned q)
{
    unsigned sum;
    asm("{\n\tadd.cc.u32 %0, %2, %3;\n\taddc.u32 %1, %1, %1;\n\t}" : "=r"(sum), "+r"(mask) : "r"(old), "r"(q));
    (void)sum;
}
// ... with a compile-time byte offset folded into the address operand
template <int OFF>
__device__ __forceinline__ unsigned atoms_add_off(unsigned addr, unsigned v)
{
    unsigned old;
    asm volatile("atom.shared.add.u32 %0, [%1+%3], %2;" : "=r"(old) : "r"(addr), "r"(v), "n"(OFF) : "memory");
    return old;
}
template <int OFF>
__device__ __forceinline__ void atoms_inc_off(unsigned addr)
{
    asm volatile("red.shared.add.u32 [%0+%1], 1;" ::"r"(addr), "n"(OFF) : "memory");
}

// shared-memory image of k_descriptor2 (one struct = one base register, see FaceSh)
#define D2_HSTRIDE (12 * 125 + 4)
struct D2Smem {
    int h_fx[2 * D2_HSTRIDE];  // lo words, then hi words, of the fixed-point histogram (5x5x5 cells)
    unsigned long long tab[32];
    FaceConst face[20];
    int lut[32];
    float kc[2];  // r2, s2
};

// Descriptor, version 2 (the default).  One CTA (8 warps) per keypoint; every warp owns every
// 8th row of the window's bounding box, 32 rows per chunk:
//   phase A  (lane = row) the x interval of the row that passes the reference's sphere /
//            descriptor-cube tests: a superset (sphere chord intersected with three slabs linear
//            in x, approximate arithmetic, widened) trimmed with the exact tests, which are weakly
//            monotone in x under IEEE rounding; the row's constants go to shared memory; a warp
//            scan of the interval lengths gives every voxel of the chunk an index;
//   phase B  (lane = a contiguous share of those indices: its consecutive voxels are x
//            neighbours, the 32 lanes sit in different rows) the per-voxel work: gradient,
//            glibc-exact window weight, rotation, icosahedron bin, 24 histogram updates.
// The histogram is FIXED POINT: integer addition is associative, so the descriptor is
// bit-reproducible from run to run (and between a tiled and a whole-volume run), and only
// native 32-bit shared atomics are needed (f32 / 64-bit shared atomics are CAS loops in SASS).
// A contribution c * 2^S (S chosen per keypoint so that |c * 2^S| < 2^31) is rounded to an
// int32 q and added to a lo/hi word pair:
//   FX_CARRY = 1: lo += q (mod 2^32) with ONE ATOMS whose returned old value gives the carry.
//     Non-negative q (the common case): straight-line code over the 8 corners of a 5x5x5-padded
//     cell grid (immediate address offsets), the 24 carry-outs collected in a mask and applied
//     to the hi words after the voxel.  Signed q: hi += sign(q) + carry, a compact loop.
//   FX_CARRY = 0: two fire-and-forget ATOMS, lo += q & 0xffff and hi += q >> 16 (arithmetic
//     shift, so q == hi * 65536 + lo for negative q too); needs < 32768 contributions per bin,
//     checked per keypoint from the size of a cell's support.  Fewer instructions but twice the
//     shared-memory wavefronts, which is what bounds this kernel.
// Windows too large for either take the 2^-32 / explicit 64-bit carry path.
#define DESC2_THREADS 256
#ifndef FX_CARRY
#define FX_CARRY 1
#endif

// OCC: resident CTAs per SM the register budget is compiled for (4: 64 registers, the
// per-voxel geometry is partly rematerialised; 3: 80 registers).  HOOK: the test hook that
// forces one of the fixed-point paths is compiled in (production instantiation: false).
template <int OCC, bool HOOK>
__global__ void __launch_bounds__(DESC2_THREADS, OCC)
    k_descriptor2(const s3d_keypoint *__restrict__ kps, int n, PyrTable T,
                  const MeshDev *__restrict__ M, unsigned char *__restrict__ out, int icos_fast)
{
    __shared__ int4 s_rows[DESC2_THREADS / 32][33];  // per warp: {first index, xa, y, z} of 32 rows
    // per warp and row: the row constants of phase B -- {py, pz, ry[0..2], rz[0..2], rowoff lo/hi}
    // -- formed in phase A by the row's own lane (all lanes busy) instead of by the one lane
    // that crosses into the row in phase B while the other 31 wait
    __shared__ float4 s_rowc[DESC2_THREADS / 32][32][3];
    __shared__ float hist[S3D_DESC_NUMEL];
    // lo and hi words of the fixed-point histogram in ONE array, so that both atomics of an
    // update share an address register.  The 4x4x4 grid of cells is stored as 5x5x5: a trilinear
    // corner that steps out of the grid (base index 3, step up) lands in a padding cell that is
    // never read back, so the scatter needs no bounds logic and every corner is the base cell's
    // address plus a COMPILE-TIME offset (folded into the ATOMS address operand).
    constexpr int HSTRIDE = D2_HSTRIDE;
    // Everything the per-voxel code touches sits in ONE struct addressed from one opaque base
    // register `sb` (see FaceSh): histogram words, expf table, face constants, lut, constants.
    __shared__ D2Smem S;
    int *const h_fx = S.h_fx;
    unsigned *h_lo = reinterpret_cast<unsigned *>(h_fx);
    int *h_hi = h_fx + HSTRIDE;
    unsigned sb = (unsigned)__cvta_generic_to_shared(&S);
    asm volatile("" : "+r"(sb));  // opaque: keeps the compiler from rematerialising it
    const unsigned h_addr = sb + (unsigned)offsetof(D2Smem, h_fx);
    unsigned long long *const s_tab = S.tab;
    FaceConst *const s_face = S.face;
    int *const s_lut = S.lut;
    float *const s_kc = S.kc;
    const FaceSh faces{sb + (unsigned)offsetof(D2Smem, face), sb + (unsigned)offsetof(D2Smem, lut)};
    const TabSh etab{sb + (unsigned)offsetof(D2Smem, tab)};
    const unsigned kc_addr = sb + (unsigned)offsetof(D2Smem, kc);
    __shared__ double s_red[DESC2_THREADS / 32];
    __shared__ float s_norm_inv;
    // S.kc: block constants the per-voxel code reads back with an opaque load: r2 and s2 derive
    // from kp.sd through f64 arithmetic, and under the register cap the compiler otherwise
    // REMATERIALISES that chain (2 DMUL + 3 F2F) for every voxel
    const int ki = blockIdx.x;
    if (ki >= n) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    s3d_keypoint kp = kps[ki];
    const int lv = kp.o * T.nlev_g + (kp.s - T.first_level);
    const float z_global = kp.z;
    kp.z = local_z(T, lv, kp.z);
    const float *__restrict__ im = T.ptrs[lv];
    const float4 *__restrict__ gim = T.gptrs ? T.gptrs[lv] : nullptr;
    const int nx = T.dims[3 * lv], ny = T.dims[3 * lv + 1], nz = T.dims[3 * lv + 2];
    const float uxf = T.units[3 * lv], uyf = T.units[3 * lv + 1], uzf = T.units[3 * lv + 2];
    const float iux = __fdiv_rn(1.0f, uxf), iuy = __fdiv_rn(1.0f, uyf), iuz = __fdiv_rn(1.0f, uzf);
    const float sigma = (float)__dmul_rn(kp.sd, 7.071067812);  // sift.c:1845-1850
    const float win_radius = (float)__dmul_rn(2.0, (double)sigma);
    const float half = (float)__ddiv_rn((double)win_radius, sqrt(2.0));
    const float desc_width = fm(2.0f, half);
    const float hist_width = __fdiv_rn(desc_width, 4.0f);
    const float bin_fctr = __fdiv_rn(1.0f, hist_width);
    const float r2 = fm(win_radius, win_radius);
    const float s2 = fm(sigma, sigma);
    int x0, x1, y0, y1, z0, z1;
    sphere_bounds_f(kp.x, win_radius, uxf, nx, x0, x1);
    sphere_bounds_f(kp.y, win_radius, uyf, ny, y0, y1);
    sphere_bounds_f(kp.z, win_radius, uzf, nz, z0, z1);
    float Rt[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) Rt[3 * i + j] = kp.R[3 * j + i];

    for (int i = tid; i < 2 * HSTRIDE; i += DESC2_THREADS) h_fx[i] = 0;
    // Fixed-point mode of this keypoint (block-uniform).  |level| <= 1 after im_scale
    // (imutil.c:1977; the blurs are convex combinations), so a gradient component is at most
    // 1/unit and |contribution| <= mag_max = sqrt(iux^2 + iuy^2 + iuz^2): S = the largest
    // power of two with mag_max * 2^S < 2^31.  A bin's support is a cube of side
    // 2 * hist_width: it must hold < 32768 voxels for the 16-bit halves not to overflow.
    float fx_scale;
    bool split;
    {
        const float mag_max = sqrtf(iux * iux + iuy * iuy + iuz * iuz) * 1.01f;
        int ex;
        frexpf(mag_max, &ex);  // mag_max < 2^ex
        fx_scale = ldexpf(1.0f, 31 - ex);
        const float support = 8.0f * hist_width * hist_width * hist_width * iux * iuy * iuz;
        split = ex <= 31 && ex >= -60 && (FX_CARRY || support * 1.25f + 512.0f < 32768.0f);
    }
    // test hook (s3d_set_option "desc_path"): 1 = signed general path, 2 = large-contribution
    // path, 3 = legacy 2^-32 / 64-bit carry path -- all must agree with the default
    const int force_path = HOOK ? (icos_fast >> 1) & 3 : 0;
    const bool trim = !(HOOK && (icos_fast & 8));  // test hook: leave the row intervals untrimmed
    // Lane-dependent order of the three vertex updates of a corner: lanes whose voxels fall into
    // the same cell and face (neighbouring rows of a smooth volume) would otherwise hit the SAME
    // address in every one of the 24 ATOMS and serialise; rotated, they hit three different bins.
    // (bit 4 of the flags switches the rotation off: A/B measurements.)
    const int rot = (icos_fast & 16) ? 0 : lane % 3;
    const unsigned rk0 = 4u * (unsigned)(13 + rot), rk1 = 4u * (unsigned)(13 + (rot + 1) % 3),
                   rk2 = 4u * (unsigned)(13 + (rot + 2) % 3);
    icos_fast &= 1;
    if (force_path == 3) split = false;
    if (tid < 32) s_tab[tid] = c_exp2f_tab[tid];
    if (tid == 0) s_kc[0] = r2, s_kc[1] = s2;
    load_faces(s_face, s_lut, M);
    __syncthreads();

    const int bx = max(x1 - x0 + 1, 0), by = max(y1 - y0 + 1, 0), bz = max(z1 - z0 + 1, 0);
    const int nrows = by * bz;
    const size_t ys = nx, zs = (size_t)nx * ny;

    // Every warp works alone until the final reduction: it owns the rows w, w+8, w+16, ... of
    // the window box (interleaved for balance), 32 rows per chunk.
    //   phase A (lane = row): the voxels of a row that pass the reference's sphere and
    //     descriptor-cube tests form ONE x interval -- a superset from the sphere chord
    //     intersected with the three slabs 0 <= vb[a] < 4 (each linear in x) in approximate
    //     arithmetic, trimmed to the exact interval with the exact tests; the row's constants
    //     are left in shared memory;
    //   phase B (lane = a contiguous share of the chunk's voxels, so that its consecutive voxels
    //     are x neighbours while the 32 lanes sit in different rows): the per-voxel work in the
    //     reference's f32 operation order (the exact tests are kept as guards; nothing fails them).
    // Only __syncwarp between the phases; the static split makes the whole computation
    // deterministic (and the fixed-point histogram makes it order-independent anyway).
    constexpr int NWARP = DESC2_THREADS / 32;
    int4 *rows = s_rows[warp];
    const int my_rows = bx > 0 && nrows > warp ? (nrows - warp + NWARP - 1) / NWARP : 0;
    // slab a: vb = (sl[a] * dxv + off_row[a]) with dxv = x - kp.x in voxels
    float sl[3];
#pragma unroll
    for (int a = 0; a < 3; a++) sl[a] = Rt[3 * a] * uxf * bin_fctr;
    for (int rc = 0; rc < my_rows; rc += 32) {
        // ---------------- phase A: one row per lane ------------------------------------------
        int cnt = 0, xa = 0, yy = 0, zz = 0;
        if (rc + lane < my_rows) {
            const int row = warp + NWARP * (rc + lane);
            yy = y0 + row % by;
            zz = z0 + row / by;
            const float vy = ((float)yy - kp.y) * uyf, vz = ((float)zz - kp.z) * uzf;
            const float rem = r2 - (vy * vy + vz * vz);
            if (rem >= -1e-3f * r2) {
                const float hx = sqrtf(fmaxf(rem, 0.0f)) * iux;
                float lo = -hx, hi = hx;  // in voxels relative to kp.x
                bool empty = false;
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    const float off = (Rt[3 * a + 1] * vy + Rt[3 * a + 2] * vz + half) * bin_fctr;
                    if (fabsf(sl[a]) < 1e-6f) {  // the slab does not depend on x
                        empty = empty || off < -1e-3f || off > 4.001f;
                    } else {
                        const float t0 = (0.0f - off) / sl[a], t1 = (4.0f - off) / sl[a];
                        lo = fmaxf(lo, fminf(t0, t1));
                        hi = fminf(hi, fmaxf(t0, t1));
                    }
                }
                if (!empty && lo <= hi + 1.0f) {
                    // widen by one voxel each side: covers every rounding of the estimates above
                    xa = max(x0, (int)floorf(kp.x + lo) - 1);
                    const int xb = min(x1, (int)ceilf(kp.x + hi) + 1);
                    cnt = max(xb - xa + 1, 0);
                }
            }
        }
        // Row constants (the SAME products the per-voxel expression forms, once per row), and
        // the interval TRIMMED with the exact tests: every test is weakly monotone in x under
        // IEEE rounding, so the voxels that pass form one interval and stepping inwards from
        // either end to the first voxel that passes yields exactly that interval.  Phase B then
        // rejects nothing on geometry (lanes stay converged); the widening above only has to
        // guarantee a superset.
        if (cnt > 0) {
            const float evy = fm(fs((float)yy, kp.y), uyf), evz = fm(fs((float)zz, kp.z), uzf);
            const float py = fm(evy, evy), pz = fm(evz, evz);
            const float ry0 = fm(Rt[1], evy), ry1 = fm(Rt[4], evy), ry2 = fm(Rt[7], evy);
            const float rz0 = fm(Rt[2], evz), rz1 = fm(Rt[5], evz), rz2 = fm(Rt[8], evz);
            auto pass = [&](int x) -> bool {
                const float vx = fm(fs((float)x, kp.x), uxf);
                if (fa(fa(fm(vx, vx), py), pz) > r2) return false;
                const float b0 = fm(fa(fa(fa(fm(Rt[0], vx), ry0), rz0), half), bin_fctr);
                const float b1 = fm(fa(fa(fa(fm(Rt[3], vx), ry1), rz1), half), bin_fctr);
                const float b2 = fm(fa(fa(fa(fm(Rt[6], vx), ry2), rz2), half), bin_fctr);
                return !(b0 < 0.0f || b0 >= 4.0f) && !(b1 < 0.0f || b1 >= 4.0f) &&
                       !(b2 < 0.0f || b2 >= 4.0f);
            };
            if (trim) {
                int xb = xa + cnt - 1;
                while (xa <= xb && !pass(xa)) xa++;
                while (xb > xa && !pass(xb)) xb--;
                cnt = max(xb - xa + 1, 0);
            }
            const size_t ro = (size_t)yy * ys + (size_t)zz * zs;
            float4 *rw = s_rowc[warp][lane];
            rw[0] = make_float4(py, pz, ry0, ry1);
            rw[1] = make_float4(ry2, rz0, rz1, rz2);
            rw[2] = make_float4(__uint_as_float((unsigned)ro), __uint_as_float((unsigned)(ro >> 32)),
                                0.0f, 0.0f);
        }
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        rows[lane] = make_int4(incl - cnt, xa, yy, zz);
        if (lane == 0) rows[32] = make_int4(total, 0, 0, 0);
        __syncwarp();
        // ---------------- phase B: per-voxel work ---------------------------------------------
        const int L = (total + 31) >> 5;
        int idx = lane * L;
        const int idx_end = min(idx + L, total);
        int r = 0;
        if (idx < idx_end) {  // last row whose first index is <= idx (empty rows share theirs)
#pragma unroll
            for (int st = 16; st > 0; st >>= 1)
                if (rows[r + st].x <= idx) r += st;
        }
        int4 cur = rows[r];
        int row_end = rows[r + 1].x;
        int x = cur.y + (idx - cur.x);
        // Row constants (y and z are fixed along a row): the y/z terms of the squared distance
        // and of the three rotated coordinates -- the SAME products the per-voxel expression
        // forms, only formed once per row -- and the row's base offset.
        float py, pz, ry[3], rz[3];
        size_t rowoff;
        auto set_row = [&](int ri) {
            const float4 *rc = s_rowc[warp][ri];
            const float4 c0 = rc[0], c1 = rc[1], c2 = rc[2];
            py = c0.x, pz = c0.y;
            ry[0] = c0.z, ry[1] = c0.w, ry[2] = c1.x;
            rz[0] = c1.y, rz[1] = c1.z, rz[2] = c1.w;
            rowoff = (size_t)__float_as_uint(c2.x) | ((size_t)__float_as_uint(c2.y) << 32);
        };
        if (idx < idx_end) set_row(r);
        for (int k = 0; k < L; k++, idx++, x++) {
            if (idx >= idx_end) break;
            if (idx >= row_end) {  // next non-empty row
                do {
                    r++;
                    row_end = rows[r + 1].x;
                } while (idx >= row_end);
                cur = rows[r];
                x = cur.y;
                set_row(r);
            }
            // geom(x, y, z) with the row terms hoisted
            const float vx = fm(fs((float)x, kp.x), uxf);
            const float sq = fa(fa(fm(vx, vx), py), pz);
            if (sq > ld_shared_f32(kc_addr)) continue;  // r2
            float vb[3];
            bool inside = true;
#pragma unroll
            for (int a = 0; a < 3; a++) {
                const float vk = fa(fa(fm(Rt[3 * a], vx), ry[a]), rz[a]);
                vb[a] = fm(fa(vk, half), bin_fctr);
                inside = inside && !(vb[a] < 0.0f || vb[a] >= 4.0f);
            }
            if (!inside) continue;
            const size_t voff = rowoff + x;
            float g[3];
            if (gim) {  // block-uniform: one 16-byte gather instead of six 4-byte ones
                const float4 g4 = __ldg(gim + voff);
                g[0] = g4.x, g[1] = g4.y, g[2] = g4.z;
            } else {
                const float *p = im + voff;
                g[0] = fm(fm(0.5f, fs(__ldg(p + 1), __ldg(p - 1))), iux);
                g[1] = fm(fm(0.5f, fs(__ldg(p + ys), __ldg(p - ys))), iuy);
                g[2] = fm(fm(0.5f, fs(__ldg(p + zs), __ldg(p - zs))), iuz);
            }
            // sift.c:1890: expf(-0.5f * sq_dist / (sigma * sigma)), f32 argument
            const float wgt_win =
                expf_glibc_t(__fdiv_rn(fm(-0.5f, sq), ld_shared_f32(kc_addr + 4u)), etab);  // / s2
            g[0] = fm(g[0], wgt_win);
            g[1] = fm(g[1], wgt_win);
            g[2] = fm(g[2], wgt_win);
            float gr[3];
#pragma unroll
            for (int a = 0; a < 3; a++)
                gr[a] = dot3(Rt[3 * a], g[0], Rt[3 * a + 1], g[1], Rt[3 * a + 2], g[2]);
            float bary[3];
            const int bin = icos_bin_t(faces, gr, bary, icos_fast != 0);
            if (bin < 0) continue;
            const float mag =
                __fsqrt_rn(fa(fa(fm(gr[0], gr[0]), fm(gr[1], gr[1])), fm(gr[2], gr[2])));
            float dv[3];
            int ib[3];
#pragma unroll
            for (int a = 0; a < 3; a++) {
                dv[a] = fs(vb[a], floorf(vb[a]));
                ib[a] = (int)vb[a];
            }
            const int i0 = faces.word(bin, rk0), i1 = faces.word(bin, rk1), i2 = faces.word(bin, rk2);
            {  // the barycentric weights in the same rotated order
                const float t0 = bary[0], t1 = bary[1], t2 = bary[2];
                bary[0] = rot == 0 ? t0 : (rot == 1 ? t1 : t2);
                bary[1] = rot == 0 ? t1 : (rot == 1 ? t2 : t0);
                bary[2] = rot == 0 ? t2 : (rot == 1 ? t0 : t1);
            }
            // mag * 2^S: scaling by a power of two commutes with every rounding below, so
            // fm(fm(mag_s, wgt), bary) == fm(fm(mag, wgt), bary) * 2^S (sift.c:1763-1765)
            const float mag_s = fm(mag, split ? fx_scale : 4294967296.0f);
            const bool small = mag_s < 2147480000.0f && force_path != 2;  // |contribution| < 2^31
            // trilinear weights (1-dx | dx)(1-dy | dy)(1-dz | dz), products associated as in
            // the reference: (x * y) * z
            const float wx0 = fs(1.0f, dv[0]), wy0 = fs(1.0f, dv[1]), wz0 = fs(1.0f, dv[2]);
            const float wxy[4] = {fm(wx0, wy0), fm(wx0, dv[1]), fm(dv[0], wy0), fm(dv[0], dv[1])};
            const int cbase = 12 * (ib[0] + 5 * ib[1] + 25 * ib[2]);  // base cell, padded layout
            const bool fast = split && small && FX_CARRY && force_path == 0 && bary[0] >= 0.0f &&
                              bary[1] >= 0.0f && bary[2] >= 0.0f;
            if (fast) {
                // Non-negative contributions (all but gradients within bary_eps outside an
                // edge): lo word += q with ONE ATOMS whose returned old value gives the carry
                // (q > ~old); hi word += 1 only then.  The three updates of a corner are issued
                // together so that their round trips overlap.
                const unsigned va0 = h_addr + 4u * (unsigned)(cbase + i0);
                const unsigned va1 = h_addr + 4u * (unsigned)(cbase + i1);
                const unsigned va2 = h_addr + 4u * (unsigned)(cbase + i2);
#define S3D_CORNER(DX, DY, DZ)                                                                \
    {                                                                                         \
        constexpr int COFF = 48 * ((DX) + 5 * (DY) + 25 * (DZ));                              \
        const float mw = fm(mag_s, fm(wxy[2 * (DX) + (DY)], (DZ) ? dv[2] : wz0));             \
        const unsigned q0 = __float2uint_rn(fm(mw, bary[0])),                                 \
                       q1 = __float2uint_rn(fm(mw, bary[1])),                                 \
                       q2 = __float2uint_rn(fm(mw, bary[2]));                                 \
        const unsigned o0 = atoms_add_off<COFF>(va0, q0), o1 = atoms_add_off<COFF>(va1, q1),  \
                       o2 = atoms_add_off<COFF>(va2, q2);                                     \
        carry_push(cmask, o0, q0);                                                            \
        carry_push(cmask, o1, q1);                                                            \
        carry_push(cmask, o2, q2);                                                            \
    }
                // cmask collects the carry-out of old + q of the 24 updates (first update =
                // bit 23): two instructions per update (IADD3 with carry-out, IADD3.X), and ONE
                // test per voxel instead of a compare/branch per corner
                unsigned cmask = 0;
                S3D_CORNER(0, 0, 0)
                S3D_CORNER(0, 0, 1)
                S3D_CORNER(0, 1, 0)
                S3D_CORNER(0, 1, 1)
                S3D_CORNER(1, 0, 0)
                S3D_CORNER(1, 0, 1)
                S3D_CORNER(1, 1, 0)
                S3D_CORNER(1, 1, 1)
                while (cmask) {  // rare: hi word += 1 for every update that wrapped its lo word
                    const int b = 31 - __clz(cmask);
                    cmask &= ~(1u << b);
                    const int u = 23 - b, c = u / 3, j = u - 3 * c;
                    const unsigned va = j == 0 ? va0 : (j == 1 ? va1 : va2);
                    atoms_inc(va + 4u * HSTRIDE +
                              48u * (unsigned)((c >> 2) + 5 * ((c >> 1) & 1) + 25 * (c & 1)));
                }
#undef S3D_CORNER
                continue;
            }
            // rare paths (a negative barycentric weight, huge contributions, test hook): one
            // compact loop over the corners, corner indices at run time
#pragma unroll 1
            for (int c = 0; c < 8; c++) {
                const int dx = c >> 2, dy = (c >> 1) & 1, dz = c & 1;
                int *const cell = h_fx + cbase + 12 * (dx + 5 * dy + 25 * dz);
                // same products as wxy[] above (no dynamically indexed array: no stack frame)
                const float wgt = fm(fm(dx ? dv[0] : wx0, dy ? dv[1] : wy0), dz ? dv[2] : wz0);
                const float mw = fm(mag_s, wgt);
                if (split && small) {
                    if (FX_CARRY) {  // signed: hi += sign(q) + carry
                        int q[3], qh[3];
                        unsigned old[3];
                        int *bp[3];
#pragma unroll
                        for (int j = 0; j < 3; j++) {
                            bp[j] = cell + (j == 0 ? i0 : (j == 1 ? i1 : i2));
                            q[j] = __float2int_rn(fm(mw, bary[j]));
                        }
#pragma unroll
                        for (int j = 0; j < 3; j++)
                            old[j] = atomicAdd(reinterpret_cast<unsigned *>(bp[j]), (unsigned)q[j]);
#pragma unroll
                        for (int j = 0; j < 3; j++)
                            qh[j] = (q[j] >> 31) + ((old[j] + (unsigned)q[j]) < old[j] ? 1 : 0);
#pragma unroll
                        for (int j = 0; j < 3; j++)
                            if (qh[j]) atomicAdd(bp[j] + HSTRIDE, qh[j]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 3; j++) {
                            int *b = cell + (j == 0 ? i0 : (j == 1 ? i1 : i2));
                            const int q = __float2int_rn(fm(mw, bary[j]));
                            atomicAdd(b, q & 0xffff);
                            atomicAdd(b + HSTRIDE, q >> 16);
                        }
                    }
                } else if (split) {  // a contribution of 2^31 units or more (never for |image| <= 1)
#pragma unroll
                    for (int j = 0; j < 3; j++) {
                        int *b = cell + (j == 0 ? i0 : (j == 1 ? i1 : i2));
                        const float cj = rintf(fm(mw, bary[j]));  // integer-valued
                        if (FX_CARRY) {
                            const long long q = __float2ll_rn(cj);
                            const unsigned ql = (unsigned)q;
                            const unsigned old = atomicAdd(reinterpret_cast<unsigned *>(b), ql);
                            const int qh = (int)(q >> 32) + ((old + ql) < old ? 1 : 0);
                            if (qh) atomicAdd(b + HSTRIDE, qh);
                        } else {
                            const float hi = floorf(cj * (1.0f / 65536.0f));  // exact
                            atomicAdd(b, __float2int_rn(cj - hi * 65536.0f));
                            atomicAdd(b + HSTRIDE, __float2int_rn(hi));
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 3; j++) {
                        int *b = cell + (j == 0 ? i0 : (j == 1 ? i1 : i2));
                        const long long q = __float2ll_rn(fm(mw, bary[j]));
                        const unsigned ql = (unsigned)q;
                        const unsigned old = atomicAdd(reinterpret_cast<unsigned *>(b), ql);
                        const int qh = (int)(q >> 32) + ((old + ql) < old ? 1 : 0);
                        if (qh) atomicAdd(b + HSTRIDE, qh);  // carry / sign word: rare
                    }
                }
            }
        }
        __syncwarp();
    }
    __syncthreads();

    // fixed point -> f32 (interior cells of the padded layout)
    for (int i = tid; i < S3D_DESC_NUMEL; i += DESC2_THREADS) {
        const int cellv = i / 12, v12 = i - 12 * cellv;
        const int p = 12 * ((cellv & 3) + 5 * ((cellv >> 2) & 3) + 25 * (cellv >> 4)) + v12;
        if (split && !FX_CARRY) {
            const long long v = (long long)h_hi[p] * 65536ll + (long long)(int)h_lo[p];
            hist[i] = (float)((double)v * (1.0 / (double)fx_scale));
        } else {
            const long long v = ((long long)h_hi[p] << 32) + (long long)h_lo[p];
            hist[i] = (float)((double)v * (1.0 / (split ? (double)fx_scale : 4294967296.0)));
        }
    }
    __syncthreads();

    // normalize_desc (sift.c:1794-1821), truncate (sift.c:1909-1915), normalize again
    const float trunc = (float)((double)(0.2f * 128.0f / S3D_DESC_NUMEL));
    for (int pass = 0; pass < 2; pass++) {
        double acc = 0.0;
        for (int i = tid; i < S3D_DESC_NUMEL; i += DESC2_THREADS) {
            const double v = (double)hist[i];
            acc = __dadd_rn(acc, __dmul_rn(v, v));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) s_red[warp] = acc;
        __syncthreads();
        if (tid == 0) {
            double tot = 0.0;
            for (int i = 0; i < DESC2_THREADS / 32; i++) tot += s_red[i];
            s_norm_inv = (float)(1.0 / (sqrt(tot) + DBL_EPSILON));
        }
        __syncthreads();
        const float ninv = s_norm_inv;
        for (int i = tid; i < S3D_DESC_NUMEL; i += DESC2_THREADS) {
            float v = fm(hist[i], ninv);
            if (pass == 0) v = fminf(v, trunc);
            hist[i] = v;
        }
        __syncthreads();
    }
    float *o32 = reinterpret_cast<float *>(out + (size_t)ki * S3D_DESC_STRIDE);
    for (int i = tid; i < S3D_DESC_NUMEL; i += DESC2_THREADS) o32[i] = hist[i];
    if (tid == 0) {
        double *o64 = reinterpret_cast<double *>(out + (size_t)ki * S3D_DESC_STRIDE +
                                                 S3D_DESC_NUMEL * sizeof(float));
        const double f = ldexp(1.0, kp.o);  // sift.c:1851, 1922-1925
        o64[0] = (double)kp.x * f;
        o64[1] = (double)kp.y * f;
        o64[2] = (double)z_global * f;
        o64[3] = kp.sd;
    }
}

// ---------------------------------------------------------------- sparse descriptor, version 3
// "Cell-owner lanes" (the default).  One CTA (8 warps) per keypoint as before, but the window is
// enumerated BY BASE CELL of the 4x4x4 descriptor grid instead of in raster order
// (desc_cell_geom.cuh): lane l of a warp owns the base cell (l & 3, (l >> 2) & 3,
// 2 * (warp & 1) + (l >> 4)) for the whole kernel and walks the voxels whose bin coordinate
// vb = (R^T v + half) * bin_fctr has floor(vb) equal to that cell; warp w takes the rows
// r = (w >> 1) (mod 4) of every cell's bounding box.  Consequences:
//   * the 32 lanes of a warp scatter into 32 DIFFERENT cells at every instruction.  The
//     histogram is stored vertex-major, h[vertex][cell] with cell (ix, iy, iz) of the padded
//     5x5x5 grid at word ix + 8 iy + 68 iz and 352 (= 0 mod 32) words per vertex plane, so the
//     bank of an update is (ix + 8 iy + 4 iz + corner offset) mod 32 whatever its vertex:
//     the 32 lanes hit 32 different banks -- one shared-memory wavefront per atomic instead of
//     the 3.6 of lanes that sit in random cells (k_descriptor2, ncu);
//   * floor(vb) is a lane constant: no floorf / F2I / cell address arithmetic per voxel, the
//     trilinear fractions are vb - ib, and the eight corner addresses are the lane's cell
//     address plus compile-time offsets;
//   * no carry logic: the histogram is one u32 per bin.  The fixed-point scale 2^S is chosen
//     per keypoint from a BOUND on the gradient magnitude inside the window (maxima over
//     8x8x1 voxel blocks, written by k_gradient next to the gradient volume) and on the number
//     of voxels in a bin's support, so that a bin cannot overflow; contributions keep >= 17
//     bits below the window's largest gradient (rounding noise of a bin ~1e-6 relative, against
//     the 1e-4 budget), and integer accumulation stays order-independent: descriptors are
//     bit-reproducible and identical between whole-volume and Z-slab-tiled runs (the block
//     maxima are per plane, so both see the same bound).
// Rows are found by a per-lane scan (d3_scan_row: ~60 instructions per row, a superset
// interval) done in rounds of four rows whenever some lane runs dry; the intervals wait in a
// small per-lane ring in shared memory, so lanes with short and long rows stay busy together.
// Every voxel passes the exact test (d3_member) before it is used.
#define D3_THREADS 256
#define D3_VSTRIDE 352
#define D3_RING 8
#define D3_WTAB 1312
struct D3Smem {
    unsigned h[12 * D3_VSTRIDE];
    unsigned long long tab[32];
    float4 face[20][5];  // FaceConst padded to 80 bytes; words 13-15 = byte offsets of the vertex planes
    int lut[32];
    float kc[4];  // r2, s2
    D3Scan scan;  // row-scan constants of the keypoint
    // window weights by squared integer distance (see the kernel): n = |offset|^2 in voxels
    float wtab[D3_WTAB];
    int4 cur[D3_THREADS];  // per-lane scan cursor {next row, rows in the bounding box, BY, ylo | zlo << 16}
    uint2 ring[D3_THREADS / 32][D3_RING][32];  // per-lane queue of scanned rows {xa | cnt << 16, y | z << 16}
};

template <int OFF>
__device__ __forceinline__ void red_add_off(unsigned addr, unsigned v)
{
    asm volatile("red.shared.add.u32 [%0+%1], %2;" ::"r"(addr), "n"(OFF), "r"(v) : "memory");
}
__device__ __forceinline__ uint2 lds_u2(unsigned addr)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u2(unsigned addr, unsigned x, unsigned y)
{
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(x), "r"(y) : "memory");
}

// packed f32x2 arithmetic (Blackwell FFMA2 / FMUL2: two results per issue slot); used only where
// the exact rounding sequence of the reference is not needed
typedef unsigned long long u64x;
__device__ __forceinline__ u64x pk2(float lo, float hi)
{
    u64x r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(u64x v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ void upk2u(u64x v, unsigned &lo, unsigned &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
}
__device__ __forceinline__ u64x fma2x(u64x a, u64x b, u64x c)
{
    u64x d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ u64x mul2x(u64x a, u64x b)
{
    u64x d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// 1.5 * 2^23: fma(a, b, MAGIC) holds round(a * b) in its low mantissa bits for |a * b| < 2^22
#define D3_MAGIC 12582912.0f
#define D3_MAGIC_BITS 0x4B400000u

template <int OCC, bool FAST, bool PRE>
__global__ void __launch_bounds__(D3_THREADS, OCC)
    k_descriptor3(const s3d_keypoint *__restrict__ kps, int n, PyrTable T,
                  const MeshDev *__restrict__ M, unsigned char *__restrict__ out)
{
    __shared__ D3Smem S;
    __shared__ float hist[S3D_DESC_NUMEL];
    __shared__ double s_red[D3_THREADS / 32];
    __shared__ float s_norm_inv;
    __shared__ float s_gmax[D3_THREADS / 32];
    unsigned sb = (unsigned)__cvta_generic_to_shared(&S);
    asm volatile("" : "+r"(sb));  // opaque: one base register for all shared data (see k_descriptor2)
    const unsigned h_addr = sb + (unsigned)offsetof(D3Smem, h);
    const FaceSh4 faces{sb + (unsigned)offsetof(D3Smem, face), sb + (unsigned)offsetof(D3Smem, lut)};
    const TabSh etab{sb + (unsigned)offsetof(D3Smem, tab)};
    const unsigned kc_addr = sb + (unsigned)offsetof(D3Smem, kc);
    const int ki = blockIdx.x;
    if (ki >= n) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    s3d_keypoint kp = kps[ki];
    const int lv = kp.o * T.nlev_g + (kp.s - T.first_level);
    const float z_global = kp.z;
    kp.z = local_z(T, lv, kp.z);
    const float4 *__restrict__ gim = T.gptrs ? T.gptrs[lv] : nullptr;
    if (gim == nullptr) __trap();  // the launcher only picks this kernel for levels with gradient volumes
    const int nx = T.dims[3 * lv], ny = T.dims[3 * lv + 1], nz = T.dims[3 * lv + 2];
    const float uxf = T.units[3 * lv], uyf = T.units[3 * lv + 1], uzf = T.units[3 * lv + 2];
    const float iux = __fdiv_rn(1.0f, uxf), iuy = __fdiv_rn(1.0f, uyf), iuz = __fdiv_rn(1.0f, uzf);
    const float sigma = (float)__dmul_rn(kp.sd, 7.071067812);  // sift.c:1845-1850
    const float win_radius = (float)__dmul_rn(2.0, (double)sigma);
    D3Key K;
    K.kx = kp.x, K.ky = kp.y, K.kz = kp.z;
    K.ux = uxf, K.uy = uyf, K.uz = uzf;
    K.half = (float)__ddiv_rn((double)win_radius, sqrt(2.0));
    const float desc_width = fm(2.0f, K.half);
    K.hw = __fdiv_rn(desc_width, 4.0f);
    K.binf = __fdiv_rn(1.0f, K.hw);
    K.r2 = fm(win_radius, win_radius);
    const float s2 = fm(sigma, sigma);
    sphere_bounds_f(kp.x, win_radius, uxf, nx, K.x0, K.x1);
    sphere_bounds_f(kp.y, win_radius, uyf, ny, K.y0, K.y1);
    sphere_bounds_f(kp.z, win_radius, uzf, nz, K.z0, K.z1);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) K.Rt[3 * i + j] = kp.R[3 * j + i];
    d3_key_finish(K);

    for (int i = tid; i < 12 * D3_VSTRIDE; i += D3_THREADS) S.h[i] = 0;
    if (tid < 32) S.tab[tid] = c_exp2f_tab[tid];
    if (tid == 0) {
        S.kc[0] = K.r2, S.kc[1] = s2;
        d3_scan_setup(K, S.scan);
    }
    if (tid < 32) S.lut[tid] = c_face_lut[tid];
    for (int i = tid; i < 20 * 16; i += D3_THREADS) {  // 16 of the 19 words of a FaceConst
        const int f = i >> 4, k = i & 15;
        unsigned w = __ldg(reinterpret_cast<const unsigned *>(M->f + f) + k);
        if (k >= 13) w *= (unsigned)(D3_VSTRIDE * 4);  // vertex index -> byte offset of its plane
        reinterpret_cast<unsigned *>(&S.face[f][0])[k] = w;
    }
    // Bound of the gradient magnitude over the window's bounding box, from the per-plane 8x8
    // block maxima of |g|^2 behind the gradient volume.
    float m2 = 0.0f;
    {
        const int nbx = (nx + 7) >> 3, nby = (ny + 7) >> 3;
        const float *__restrict__ bm = reinterpret_cast<const float *>(gim + (size_t)nx * ny * nz);
        const int bx0 = K.x0 >> 3, by0 = K.y0 >> 3;
        const int wbx = (K.x1 >> 3) - bx0 + 1, wby = (K.y1 >> 3) - by0 + 1, wz = K.z1 - K.z0 + 1;
        const int tot = wbx > 0 && wby > 0 && wz > 0 ? wbx * wby * wz : 0;
        for (int t = tid; t < tot; t += D3_THREADS) {
            const int bxi = t % wbx, r = t / wbx;
            const int byi = r % wby, zi = r / wby;
            m2 = fmaxf(m2, __ldg(bm + ((size_t)(K.z0 + zi) * nby + by0 + byi) * nbx + bx0 + bxi));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m2 = fmaxf(m2, __shfl_xor_sync(0xffffffffu, m2, o));
        if (lane == 0) s_gmax[warp] = m2;
    }
    // Window weights from a table.  For an integer centre and isotropic power-of-two units u
    // (every detector keypoint of a volume with such units) the offsets v are u * integers, so
    // the squared distance the reference forms, (vx*vx + vy*vy) + vz*vz in f32, is EXACTLY
    // u^2 * n with n = dx^2 + dy^2 + dz^2 < 2^24: the weight expf(-0.5f * sq / sigma^2)
    // (sift.c:1890) depends on n alone.  Tabulated once per keypoint with the per-voxel
    // arithmetic (bit-identical), it replaces an IEEE division and the f64 exp (33 of ~380
    // instructions per voxel).  Other keypoints take the arithmetic path (block-uniform).
    bool tab_ok;
    float inv_u2 = 0.0f, nmaxf = 0.0f;
    {
        int ex;
        const bool pow2 = frexpf(uxf, &ex) == 0.5f;
        const bool iso = uxf == uyf && uyf == uzf;
        const bool integer = kp.x == floorf(kp.x) && kp.y == floorf(kp.y) && kp.z == floorf(kp.z);
        inv_u2 = iux * iux;
        nmaxf = K.r2 * inv_u2;
        tab_ok = pow2 && iso && integer && nmaxf < (float)(D3_WTAB - 1);
    }
    const unsigned wtab_addr = sb + (unsigned)offsetof(D3Smem, wtab);
    // ---- this lane's cell and its rows ---------------------------------------------------
    const int ib0 = lane & 3, ib1 = (lane >> 2) & 3, ib2 = 2 * (warp & 1) + (lane >> 4);
    bool scan_done;
    {
        D3Cell Cc;
        d3_cell_bbox(K, ib0, ib1, ib2, Cc);
        const int BY = Cc.yhi - Cc.ylo + 1, BZ = Cc.zhi - Cc.zlo + 1;
        const int nrows = BY > 0 && BZ > 0 ? BY * BZ : 0;
        // rows (warp >> 1) + 4k of the bounding box, in (z, y) order
        S.cur[tid] = make_int4(warp >> 1, nrows, max(BY, 1), max(Cc.ylo, 0) | (max(Cc.zlo, 0) << 16));
        scan_done = (warp >> 1) >= nrows;
    }
    __syncthreads();
    if (tab_ok) {  // the exp table (S.tab) is visible now; the second barrier publishes wtab
        const int nmax = (int)nmaxf;
        const float u2 = uxf * uxf;
        for (int nn = tid; nn <= nmax; nn += D3_THREADS)
            S.wtab[nn] = expf_glibc_t(__fdiv_rn(fm(-0.5f, fm((float)nn, u2)), s2), etab);
    }
#pragma unroll
    for (int i = 0; i < D3_THREADS / 32; i++) m2 = fmaxf(m2, s_gmax[i]);
    // Fixed-point scale 2^S (block-uniform).  A contribution is mag * 2^S * w_c * bary with
    // mag <= gmax * 1.001 (window weight <= 1; the rotation keeps the norm to 1e-6),
    // bary <= 1 + 2 bary_eps and w_c the trilinear weight of the voxel for the bin's cell,
    // a product of three hat functions of the rotated coordinate.  Summed over the voxel
    // lattice, sum_v w_c(v) <= integral of sup_{|y - x| <= D/2} w_c(y) dx / voxel volume
    // <= hist_width^3 * (1 + 2 delta)^3 / (ux uy uz), with D the voxel diagonal and
    // delta = D / (2 hist_width): each hat is 1-Lipschitz in bin units, and
    // int min(1, 1 + delta - |t|) dt = 1 + 2 delta.  So a bin stays below
    // gmax * 2^S * 1.003 * NW + (half a unit of rounding for each of at most NS voxels in its
    // support) < 0xE0000000; a bin that reads back >= 0xF0000000 holds a (tiny) NEGATIVE sum
    // of contributions with barycentric weights in [-bary_eps, 0).  A single contribution
    // stays below 2^22 (the magic-number rounding below).
    float fx_scale = 1.0f;
    {
        const float gmax = sqrtf(m2) * 1.001f;
        const float D = sqrtf(uxf * uxf + uyf * uyf + uzf * uzf);
        const float side = 2.0f * K.hw + D;
        const float NS = side * side * side * iux * iuy * iuz * 1.01f + 64.0f;
        const float wside = K.hw + D;  // hist_width * (1 + 2 delta)
        const float NW = wside * wside * wside * iux * iuy * iuz * 1.01f + 1.0f;
        const float lim = fminf((3.7e9f - 0.5f * NS) / (NW * 1.003f), 4.0e6f);
        if (gmax > 0.0f && lim > 0.0f) {
            int ex;
            frexpf(lim / gmax, &ex);  // lim / gmax = m * 2^ex, m in [0.5, 1)
            ex = min(max(ex - 1, -100), 100);
            fx_scale = ldexpf(1.0f, ex);
        }
    }
    __syncthreads();

    float ibf[3] = {(float)ib0, (float)ib1, (float)ib2};
    asm volatile("" : "+f"(ibf[0]), "+f"(ibf[1]), "+f"(ibf[2]));  // keep, do not re-derive from tid
    const unsigned cell_addr = h_addr + 4u * (unsigned)(ib0 + 8 * ib1 + 68 * ib2);
    unsigned ring_addr = sb + (unsigned)offsetof(D3Smem, ring) + 8u * (unsigned)(warp * D3_RING * 32 + lane);
    asm volatile("" : "+r"(ring_addr));
    int wr = 0, nq = 0;   // ring write slot, queued rows
    int rem = 0;          // voxels left in the current row
    float xf = 0.0f;
    unsigned yz = 0;      // y | z << 16 of the current row
    unsigned vidx = 0;    // voxel index of the current voxel in the level
    const float mag_scale = fx_scale;
    float4 gpre = make_float4(0.f, 0.f, 0.f, 0.f);
    bool pre_ok = false;

    for (;;) {
        const bool starving = rem == 0 && nq == 0 && !scan_done;
        if (__any_sync(0xffffffffu, starving)) {
            // a round of four rows for every lane that has room in its queue
            int4 cur = S.cur[tid];
            const float inv_by = __frcp_rn((float)cur.z);
#pragma unroll 1
            for (int k = 0; k < 4; k++) {
                if (nq < D3_RING && cur.x < cur.y) {
                    const int dz = (int)(((float)cur.x + 0.5f) * inv_by);
                    const int y = (cur.w & 0xffff) + (cur.x - dz * cur.z), z = (cur.w >> 16) + dz;
                    int xa, cnt;
                    d3_scan_row(S.scan, ibf, y, z, xa, cnt);
                    if (cnt > 0) {
                        sts_u2(ring_addr + 256u * (unsigned)wr, (unsigned)xa | ((unsigned)cnt << 16),
                               (unsigned)y | ((unsigned)z << 16));
                        wr = (wr + 1) & (D3_RING - 1);
                        nq++;
                    }
                    cur.x += 4;
                }
            }
            S.cur[tid].x = cur.x;
            scan_done = cur.x >= cur.y;
            continue;
        }
        const bool active = rem > 0 || nq > 0;
        if (!__any_sync(0xffffffffu, active)) break;
        if (!active) continue;
        if (rem == 0) {  // next queued row
            const uint2 rec = lds_u2(ring_addr + 256u * (unsigned)((wr - nq) & (D3_RING - 1)));
            nq--;
            const unsigned xa = rec.x & 0xffffu;
            rem = (int)(rec.x >> 16);
            yz = rec.y;
            xf = (float)xa;
            vidx = xa + (unsigned)nx * ((yz & 0xffffu) + (unsigned)ny * (yz >> 16));
            pre_ok = false;
        }
        // ---- one voxel (sift.c:1866-1905) --------------------------------------------------
        float sq, dv[3];
        const bool member = d3_member(K, ibf, ld_shared_f32(kc_addr), xf, (float)(yz & 0xffffu),
                                      (float)(yz >> 16), sq, dv);
        const unsigned vi = vidx;
        xf = fa(xf, 1.0f);
        vidx++;
        rem--;
        // the next voxel of the row is fetched one trip ahead (PRE): the gather is the longest
        // latency of the loop and the lanes sit in 32 different lines
        float4 g4;
        if (PRE) {
            g4 = pre_ok ? gpre : __ldg(gim + vi);
            pre_ok = rem > 0;
            if (pre_ok) gpre = __ldg(gim + vidx);
        } else {
            g4 = __ldg(gim + vi);
        }
        if (!member) continue;
        float g[3] = {g4.x, g4.y, g4.z};
        // sift.c:1890: expf(-0.5f * sq_dist / (sigma * sigma)), f32 argument, glibc's expf
        float wgt_win;
        if (tab_ok)
            wgt_win = ld_shared_f32(wtab_addr + 4u * (unsigned)__float2int_rn(fm(sq, inv_u2)));
        else
            wgt_win = expf_glibc_t(__fdiv_rn(fm(-0.5f, sq), ld_shared_f32(kc_addr + 4u)), etab);
        g[0] = fm(g[0], wgt_win);
        g[1] = fm(g[1], wgt_win);
        g[2] = fm(g[2], wgt_win);
        float gr[3];
#pragma unroll
        for (int a = 0; a < 3; a++)
            gr[a] = dot3(K.Rt[3 * a], g[0], K.Rt[3 * a + 1], g[1], K.Rt[3 * a + 2], g[2]);
        float bary[3];
        int voff[3];
        const int bin = icos_bin4<FAST>(faces, gr, bary, voff);
        if (bin < 0) continue;
        const float mag = __fsqrt_rn(fa(fa(fm(gr[0], gr[0]), fm(gr[1], gr[1])), fm(gr[2], gr[2])));
        // The scatter (sift.c:1763-1765: (mag * w_c) * bary_j into the 8 cells x 3 vertices) in
        // fixed point.  The products are formed as w_c * (mag * 2^S * bary_j) in packed f32x2
        // arithmetic -- a re-association worth 1e-7 relative per contribution, far inside the
        // fixed-point rounding -- and rounded to an integer by ONE fused multiply-add with the
        // magic constant (the float-to-int conversion is a quarter-rate instruction, and there
        // are 24 per voxel): 12 FFMA2 instead of 32 FMUL + 24 F2I.
        const float mag_s = fm(mag, mag_scale);
        const unsigned va0 = cell_addr + (unsigned)voff[0];
        const unsigned va1 = cell_addr + (unsigned)voff[1];
        const unsigned va2 = cell_addr + (unsigned)voff[2];
        const float wx0 = fs(1.0f, dv[0]), wy0 = fs(1.0f, dv[1]), wz0 = fs(1.0f, dv[2]);
        const u64x mb01 = mul2x(pk2(mag_s, mag_s), pk2(bary[0], bary[1]));
        const float mb2 = fm(mag_s, bary[2]);
        const u64x mb22 = pk2(mb2, mb2);
        const u64x wz = pk2(wz0, dv[2]);
        const u64x magic2 = pk2(D3_MAGIC, D3_MAGIC);
        // (x * y) pairs, then (x * y) * z: w[2 * DX + DY] = {DZ = 0, DZ = 1}
        const u64x wxy_a = mul2x(pk2(wx0, wx0), pk2(wy0, dv[1]));      // (0,0) (0,1)
        const u64x wxy_b = mul2x(pk2(dv[0], dv[0]), pk2(wy0, dv[1]));  // (1,0) (1,1)
        float wxy00, wxy01, wxy10, wxy11;
        upk2(wxy_a, wxy00, wxy01);
        upk2(wxy_b, wxy10, wxy11);
#define S3D_PAIR3(WXY, DX, DY)                                                                     \
    {                                                                                              \
        constexpr int C0 = 4 * ((DX) + 8 * (DY)), C1 = C0 + 4 * 68;                                \
        const u64x w = mul2x(pk2(WXY, WXY), wz); /* {w(DZ=0), w(DZ=1)} */                          \
        float w0, w1;                                                                              \
        upk2(w, w0, w1);                                                                           \
        unsigned q00, q01, q10, q11, q20, q21;                                                     \
        upk2u(fma2x(pk2(w0, w0), mb01, magic2), q00, q01);                                         \
        upk2u(fma2x(pk2(w1, w1), mb01, magic2), q10, q11);                                         \
        upk2u(fma2x(w, mb22, magic2), q20, q21);                                                   \
        red_add_off<C0>(va0, q00 - D3_MAGIC_BITS);                                                 \
        red_add_off<C0>(va1, q01 - D3_MAGIC_BITS);                                                 \
        red_add_off<C0>(va2, q20 - D3_MAGIC_BITS);                                                 \
        red_add_off<C1>(va0, q10 - D3_MAGIC_BITS);                                                 \
        red_add_off<C1>(va1, q11 - D3_MAGIC_BITS);                                                 \
        red_add_off<C1>(va2, q21 - D3_MAGIC_BITS);                                                 \
    }
        S3D_PAIR3(wxy00, 0, 0)
        S3D_PAIR3(wxy01, 0, 1)
        S3D_PAIR3(wxy10, 1, 0)
        S3D_PAIR3(wxy11, 1, 1)
#undef S3D_PAIR3
    }
    __syncthreads();

    // fixed point -> f32 (interior cells of the padded layout)
    {
        const double inv = 1.0 / (double)fx_scale;
        for (int i = tid; i < S3D_DESC_NUMEL; i += D3_THREADS) {
            const int cellv = i / 12, v12 = i - 12 * cellv;
            const unsigned u = S.h[v12 * D3_VSTRIDE + (cellv & 3) + 8 * ((cellv >> 2) & 3) + 68 * (cellv >> 4)];
            const long long v = u >= 0xF0000000u ? (long long)u - 4294967296ll : (long long)u;
            hist[i] = (float)((double)v * inv);
        }
    }
    __syncthreads();

    // normalize_desc (sift.c:1794-1821), truncate (sift.c:1909-1915), normalize again
    const float trunc = (float)((double)(0.2f * 128.0f / S3D_DESC_NUMEL));
    for (int pass = 0; pass < 2; pass++) {
        double acc = 0.0;
        for (int i = tid; i < S3D_DESC_NUMEL; i += D3_THREADS) {
            const double v = (double)hist[i];
            acc = __dadd_rn(acc, __dmul_rn(v, v));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) s_red[warp] = acc;
        __syncthreads();
        if (tid == 0) {
            double tot = 0.0;
            for (int i = 0; i < D3_THREADS / 32; i++) tot += s_red[i];
            s_norm_inv = (float)(1.0 / (sqrt(tot) + DBL_EPSILON));
        }
        __syncthreads();
        const float ninv = s_norm_inv;
        for (int i = tid; i < S3D_DESC_NUMEL; i += D3_THREADS) {
            float v = fm(hist[i], ninv);
            if (pass == 0) v = fminf(v, trunc);
            hist[i] = v;
        }
        __syncthreads();
    }
    float *o32 = reinterpret_cast<float *>(out + (size_t)ki * S3D_DESC_STRIDE);
    for (int i = tid; i < S3D_DESC_NUMEL; i += D3_THREADS) o32[i] = hist[i];
    if (tid == 0) {
        double *o64 = reinterpret_cast<double *>(out + (size_t)ki * S3D_DESC_STRIDE +
                                                 S3D_DESC_NUMEL * sizeof(float));
        const double f = ldexp(1.0, kp.o);  // sift.c:1851, 1922-1925
        o64[0] = (double)kp.x * f;
        o64[1] = (double)kp.y * f;
        o64[2] = (double)z_global * f;
        o64[3] = kp.sd;
    }
}

// ---------------------------------------------------------------- dense descriptors
// extract_dense_descriptors_no_rotate (sift.c:2462-2480): barycentric weights of
// the gradient direction, written to three of the twelve channels.
__global__ void __launch_bounds__(256)
    k_dense_bary(const float *__restrict__ sm, int nx, int ny, int nz, float iux, float iuy,
                 float iuz, const MeshDev *__restrict__ M, float *__restrict__ temp)
{
    __shared__ FaceConst s_face[20];
    __shared__ int s_lut[32];
    load_faces(s_face, s_lut, M);
    __syncthreads();
    const size_t total = (size_t)nx * ny * nz;
    const size_t ys = nx, zs = (size_t)nx * ny;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(idx % nx);
        const size_t r = idx / nx;
        const int y = (int)(r % ny);
        const int z = (int)(r / ny);
        float h[12];
#pragma unroll
        for (int k = 0; k < 12; k++) h[k] = 0.0f;
        if (x >= 1 && x <= nx - 2 && y >= 1 && y <= ny - 2 && z >= 1 && z <= nz - 2) {
            const float *p = sm + idx;
            float g[3], bary[3];
            g[0] = fm(fm(0.5f, fs(__ldg(p + 1), __ldg(p - 1))), iux);
            g[1] = fm(fm(0.5f, fs(__ldg(p + ys), __ldg(p - ys))), iuy);
            g[2] = fm(fm(0.5f, fs(__ldg(p + zs), __ldg(p - zs))), iuz);
            const int bin = icos_bin(s_face, s_lut, g, bary);
            if (bin >= 0) {
#pragma unroll
                for (int k = 0; k < 12; k++) {
                    if (k == s_face[bin].idx[0]) h[k] = bary[0];
                    if (k == s_face[bin].idx[1]) h[k] = bary[1];
                    if (k == s_face[bin].idx[2]) h[k] = bary[2];
                }
            }
        }
        float4 *o = reinterpret_cast<float4 *>(temp + idx * 12);
        o[0] = make_float4(h[0], h[1], h[2], h[3]);
        o[1] = make_float4(h[4], h[5], h[6], h[7]);
        o[2] = make_float4(h[8], h[9], h[10], h[11]);
    }
}

// extract_dense_descriptors_rotate (sift.c:2521-2588): per voxel, an orientation from the
// structure tensor of a sigma0*ori_sig_fctr window (identity when rejected), then one
// 12-bin histogram of the gradients in a sphere of radius 2*desc_sigma, rotated by R^T
// (extract_dense_descrip_rotate, sift.c:2295-2343).  One thread per voxel, both windows
// walked in the reference's raster order, so every f32 sum rounds as on the CPU.
#define DROT_THREADS 128
__global__ void __launch_bounds__(DROT_THREADS)
    k_dense_rotate(const float *__restrict__ sm, int nx, int ny, int nz, float uxf, float uyf,
                   float uzf, double ori_sigma, double desc_sigma, double corner_thresh,
                   const MeshDev *__restrict__ M, float *__restrict__ out)
{
    __shared__ unsigned long long s_tab[32];
    __shared__ FaceConst s_face[20];
    __shared__ int s_lut[32];
    __shared__ float s_h[12][DROT_THREADS];
    if (threadIdx.x < 32) s_tab[threadIdx.x] = c_exp2f_tab[threadIdx.x];
    load_faces(s_face, s_lut, M);
    __syncthreads();
    const size_t total = (size_t)nx * ny * nz;
    const size_t idx = (size_t)blockIdx.x * DROT_THREADS + threadIdx.x;
    if (idx >= total) return;
    const int x = (int)(idx % nx);
    const size_t rr = idx / nx;
    const int y = (int)(rr % ny);
    const int z = (int)(rr / ny);
    const float vcx = (float)x, vcy = (float)y, vcz = (float)z;

    float R[9];
    double conf;
    const bool ok = orient_core(sm, nx, ny, nz, uxf, uyf, uzf, vcx, vcy, vcz, ori_sigma,
                                corner_thresh, s_tab, R, conf);
    float Rt[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) Rt[3 * i + j] = ok ? R[3 * j + i] : (i == j ? 1.0f : 0.0f);

    const float win_radius = (float)__dmul_rn(2.0, desc_sigma);  // desc_rad_fctr, sift.c:2306
    const float r2 = fm(win_radius, win_radius);
    const double s2 = __dmul_rn(desc_sigma, desc_sigma);
    const float iux = __fdiv_rn(1.0f, uxf), iuy = __fdiv_rn(1.0f, uyf), iuz = __fdiv_rn(1.0f, uzf);
    int x0, x1, y0, y1, z0, z1;
    sphere_bounds_f(vcx, win_radius, uxf, nx, x0, x1);
    sphere_bounds_f(vcy, win_radius, uyf, ny, y0, y1);
    sphere_bounds_f(vcz, win_radius, uzf, nz, z0, z1);
    const size_t ys = nx, zs = (size_t)nx * ny;
    float *h = &s_h[0][threadIdx.x];
#pragma unroll
    for (int k = 0; k < 12; k++) h[k * DROT_THREADS] = 0.0f;

    for (int zz = z0; zz <= z1; zz++) {
        const float dz = fm(fs((float)zz, vcz), uzf);
        const float dz2 = fm(dz, dz);
        for (int yy = y0; yy <= y1; yy++) {
            const float dy = fm(fs((float)yy, vcy), uyf);
            const float dy2 = fm(dy, dy);
            const float *row = sm + (size_t)yy * ys + (size_t)zz * zs;
            for (int xx = x0; xx <= x1; xx++) {
                const float dx = fm(fs((float)xx, vcx), uxf);
                const float sq = fa(fa(fm(dx, dx), dy2), dz2);
                if (sq > r2) continue;
                const float *p = row + xx;
                float g[3], gr[3], bary[3];
                g[0] = fm(fm(0.5f, fs(__ldg(p + 1), __ldg(p - 1))), iux);
                g[1] = fm(fm(0.5f, fs(__ldg(p + ys), __ldg(p - ys))), iuy);
                g[2] = fm(fm(0.5f, fs(__ldg(p + zs), __ldg(p - zs))), iuz);
#pragma unroll
                for (int a = 0; a < 3; a++)
                    gr[a] = dot3(Rt[3 * a], g[0], Rt[3 * a + 1], g[1], Rt[3 * a + 2], g[2]);
                const int bin = icos_bin(s_face, s_lut, gr, bary);
                if (bin < 0) continue;
                // magnitude of the unrotated gradient (sift.c:2330)
                const float mag = __fsqrt_rn(fa(fa(fm(g[0], g[0]), fm(g[1], g[1])), fm(g[2], g[2])));
                // sift.c:2333: f32 product, f64 divide by sigma^2, narrowed for expf
                const float arg = (float)__ddiv_rn((double)fm(-0.5f, sq), s2);
                const float w = expf_glibc(arg, s_tab);
                const float mw = fm(mag, w);
                float *h0 = h + s_face[bin].idx[0] * DROT_THREADS;
                float *h1 = h + s_face[bin].idx[1] * DROT_THREADS;
                float *h2 = h + s_face[bin].idx[2] * DROT_THREADS;
                *h0 = fa(*h0, fm(mw, bary[0]));
                *h1 = fa(*h1, fm(mw, bary[1]));
                *h2 = fa(*h2, fm(mw, bary[2]));
            }
        }
    }
    float4 *o = reinterpret_cast<float4 *>(out + idx * 12);
    o[0] = make_float4(h[0], h[1 * DROT_THREADS], h[2 * DROT_THREADS], h[3 * DROT_THREADS]);
    o[1] = make_float4(h[4 * DROT_THREADS], h[5 * DROT_THREADS], h[6 * DROT_THREADS],
                       h[7 * DROT_THREADS]);
    o[2] = make_float4(h[8 * DROT_THREADS], h[9 * DROT_THREADS], h[10 * DROT_THREADS],
                       h[11 * DROT_THREADS]);
}

// postproc_Hist (sift.c:2267-2292): normalise, clamp, normalise, x intensity
__global__ void __launch_bounds__(256)
    k_dense_post(float *__restrict__ desc, const float *__restrict__ raw, size_t nvox)
{
    const float hist_trunc = (float)((double)(0.2f * 128.0f / S3D_DESC_NUMEL) * S3D_DESC_NUMEL / 12);
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < nvox;
         idx += (size_t)gridDim.x * blockDim.x) {
        float4 *p = reinterpret_cast<float4 *>(desc + idx * 12);
        float4 a = p[0], b = p[1], c = p[2];
        float h[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
#pragma unroll
        for (int pass = 0; pass < 2; pass++) {
            double nrm = 0.0;
#pragma unroll
            for (int k = 0; k < 12; k++) nrm = __dadd_rn(nrm, __dmul_rn((double)h[k], (double)h[k]));
            nrm = sqrt(nrm) + DBL_EPSILON;
            const float ninv = (float)(1.0 / nrm);
#pragma unroll
            for (int k = 0; k < 12; k++) {
                h[k] = fm(h[k], ninv);
                if (pass == 0) h[k] = fminf(h[k], hist_trunc);
            }
        }
        const float val = __ldg(raw + idx);
#pragma unroll
        for (int k = 0; k < 12; k++) h[k] = fm(h[k], val);
        p[0] = make_float4(h[0], h[1], h[2], h[3]);
        p[1] = make_float4(h[4], h[5], h[6], h[7]);
        p[2] = make_float4(h[8], h[9], h[10], h[11]);
    }
}

// ---------------------------------------------------------------- gradient volumes
// SIFT3D_IM_GET_GRAD (immacros.h:105-111) scaled by 1/units (IM_GET_GRAD_ISO, sift.c:150-155):
// g = (0.5f * (v[+1] - v[-1])) * (1 / unit) per axis, f32, separately rounded -- the expression
// both assign_eig_ori (sift.c:1385) and extract_descrip (sift.c:1882) evaluate per visit.
__global__ void __launch_bounds__(256)
    k_gradient(const float *__restrict__ im, int nx, int ny, int nz, float iux, float iuy,
               float iuz, float4 *__restrict__ out, float *__restrict__ bmax, int nbx, int nby)
{   // grid: (ceil(nx / 32), ceil(ny / 8), ceil(nz / 8)), block 256: warp w owns plane
    // 8 * blockIdx.z + w and walks the 8 rows of its y block, lanes along x (512-byte stores).
    // Also leaves max |g|^2 over every 8 x 8 x 1 block of voxels in bmax[z][y/8][x/8]
    // (k_descriptor3 bounds the gradient magnitude inside a window with it): a register max
    // over the rows, then a 3-step butterfly over the 8 lanes of an x block -- no shared memory.
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x = blockIdx.x * 32 + lane, z = blockIdx.z * 8 + warp;
    if (z >= nz) return;
    const size_t ys = nx, zs = (size_t)nx * ny;
    float m2 = 0.0f;
    const bool xin = x >= 1 && x <= nx - 2, zin = z >= 1 && z <= nz - 2;
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const int y = blockIdx.y * 8 + r;
        if (x < nx && y < ny) {
            const size_t idx = x + y * ys + z * zs;
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
            if (xin && zin && y >= 1 && y <= ny - 2) {
                const float *p = im + idx;
                g.x = fm(fm(0.5f, fs(__ldg(p + 1), __ldg(p - 1))), iux);
                g.y = fm(fm(0.5f, fs(__ldg(p + ys), __ldg(p - ys))), iuy);
                g.z = fm(fm(0.5f, fs(__ldg(p + zs), __ldg(p - zs))), iuz);
            }
            out[idx] = g;
            m2 = fmaxf(m2, g.x * g.x + g.y * g.y + g.z * g.z);
        }
    }
    m2 = fmaxf(m2, __shfl_xor_sync(0xffffffffu, m2, 1));
    m2 = fmaxf(m2, __shfl_xor_sync(0xffffffffu, m2, 2));
    m2 = fmaxf(m2, __shfl_xor_sync(0xffffffffu, m2, 4));
    const int bxi = blockIdx.x * 4 + (lane >> 3);
    if ((lane & 7) == 0 && bxi < nbx) bmax[((size_t)z * nby + blockIdx.y) * nbx + bxi] = m2;
}

PyrTable make_table(const s3d_engine *e)
{
    PyrTable T;
    T.ptrs = e->d_level_ptrs;
    T.dims = e->d_level_dims;
    T.units = e->d_level_units;
    T.scales = e->d_level_scales;
    T.nlev_g = e->nlev_g;
    T.first_level = e->first_level;
    T.zoff = e->slab.empty() ? nullptr : e->d_level_zoff;
    T.gptrs = e->grad_valid ? e->d_level_gptrs : nullptr;
    return T;
}

}  // namespace

void s3d_gradients_free(s3d_engine *e)
{
    for (float4 *p : e->grad)
        if (p) cudaFree(p);
    e->grad.clear();
    e->grad_cap.clear();
    e->grad_uploaded.clear();
    if (e->d_level_gptrs) cudaFree(e->d_level_gptrs);
    e->d_level_gptrs = nullptr;
    e->grad_valid = false;
}

// Gradient volumes for the keypoint levels s = 0..K-1 of the resident pyramid.  Best effort:
// a level whose volume cannot be allocated keeps a null entry and the kernels gather scalars.
int s3d_gradients_prepare(s3d_engine *e)
{
    if (e->grad_valid) return 0;
    const int L = (int)e->g.size();
    if (L == 0) return 0;
    if ((int)e->grad.size() != L) {
        s3d_gradients_free(e);
        e->grad.assign(L, nullptr);
        e->grad_cap.assign(L, 0);
    }
    if (!e->d_level_gptrs) S3D_CUDA(e, cudaMalloc(&e->d_level_gptrs, L * sizeof(float4 *)));
    for (int lv = 0; lv < L; lv++) {
        const int sidx = lv % e->nlev_g + e->first_level;
        const LevelDev &l = e->g[lv];
        const bool want = sidx >= 0 && sidx < e->K && l.d && l.n() > 0;
        if (!want) continue;
        if (e->grad_cap[lv] < l.n()) {
            if (e->grad[lv]) cudaFree(e->grad[lv]);
            e->grad[lv] = nullptr;
            e->grad_cap[lv] = 0;
            // the gradient volume, then the per-plane 8x8 block maxima of |g|^2 (k_gradient)
            const size_t nbm = (size_t)((l.g.nx + 7) / 8) * ((l.g.ny + 7) / 8) * l.g.nz;
            if (cudaMalloc(&e->grad[lv], l.n() * sizeof(float4) + nbm * sizeof(float)) != cudaSuccess) {
                cudaGetLastError();  // out of memory: this level stays on the scalar path
                e->grad[lv] = nullptr;
                continue;
            }
            e->grad_cap[lv] = l.n();
        }
        const float ux = (float)l.g.ux, uy = (float)l.g.uy, uz = (float)l.g.uz;
        if (l.g.ny > 8 * 65535 || l.g.nz > 8 * 65535 || l.g.ny > 65535 || l.g.nz > 65535 ||
            l.n() >= ((size_t)1 << 32)) {  // grid limits, 16-bit row records, 32-bit voxel index: scalar path
            cudaFree(e->grad[lv]);
            e->grad[lv] = nullptr;
            e->grad_cap[lv] = 0;
            continue;
        }
        const int nbx = (l.g.nx + 7) / 8, nby = (l.g.ny + 7) / 8;
        k_gradient<<<dim3((l.g.nx + 31) / 32, nby, (l.g.nz + 7) / 8), 256, 0, e->stream>>>(
            l.d, l.g.nx, l.g.ny, l.g.nz, __fdiv_rn_host(ux), __fdiv_rn_host(uy), __fdiv_rn_host(uz),
            e->grad[lv], reinterpret_cast<float *>(e->grad[lv] + l.n()), nbx, nby);
        S3D_LAUNCH_CHECK(e);
    }
    if (e->grad_uploaded != e->grad) {  // the device table only changes when a volume was (re)allocated
        S3D_CUDA(e, cudaMemcpyAsync(e->d_level_gptrs, e->grad.data(), L * sizeof(float4 *),
                                    cudaMemcpyHostToDevice, e->stream));
        S3D_CUDA(e, cudaStreamSynchronize(e->stream));  // e->grad is host memory that may change
        e->grad_uploaded = e->grad;
    }
    e->grad_valid = true;
    return 0;
}

int s3d_upload_mesh(s3d_engine *e, const float *v, const int *idx)
{
    // per-face constants in the reference's f32 order (cart2bary, sift.c:343-364)
    MeshDev M;
    for (int i = 0; i < 20; i++) {
        const float *v0 = v + 9 * i, *v1 = v0 + 3, *v2 = v0 + 6;
        FaceConst &F = M.f[i];
        for (int j = 0; j < 3; j++) {
            volatile float e1 = v1[j] - v0[j];
            volatile float e2 = v2[j] - v0[j];
            volatile float t = v0[j] * -1.0f;
            F.e1[j] = e1;
            F.e2[j] = e2;
            F.t[j] = t;
            F.idx[j] = idx[3 * i + j];
        }
        {
            volatile float a, b;
            a = F.t[1] * F.e1[2];
            b = F.t[2] * F.e1[1];
            F.q[0] = a - b;
            a = F.t[2] * F.e1[0];
            b = F.t[0] * F.e1[2];
            F.q[1] = a - b;
            a = F.t[0] * F.e1[1];
            b = F.t[1] * F.e1[0];
            F.q[2] = a - b;
            volatile float s0 = F.e2[0] * F.q[0], s1 = F.e2[1] * F.q[1], s2 = F.e2[2] * F.q[2];
            volatile float s01 = s0 + s1;
            F.e2q = s01 + s2;
        }
        double c[3] = {0, 0, 0}, nrm = 0;
        for (int j = 0; j < 3; j++) c[j] = ((double)v0[j] + v1[j] + v2[j]) / 3.0;
        for (int j = 0; j < 3; j++) nrm += c[j] * c[j];
        nrm = sqrt(nrm);
        for (int j = 0; j < 3; j++) F.vmid[j] = (float)(c[j] / nrm);
    }
    for (int i = 0; i < 12; i++)
        for (int j = 0; j < 3; j++) M.vert[i][j] = 0.0f;
    {
        float vm[20][3];
        for (int i = 0; i < 20; i++)
            for (int j = 0; j < 3; j++) vm[i][j] = M.f[i].vmid[j];
        S3D_CUDA(e, cudaMemcpyToSymbol(c_vmid, vm, sizeof(vm)));
        const double ph = 1.6180339887, ip = 1.0 / ph;
        const double rep[4][3] = {{1, 1, 1}, {ip, 0, ph}, {ph, ip, 0}, {0, ph, ip}};
        int lut[32];
        for (int t = 0; t < 4; t++)
            for (int sb = 0; sb < 8; sb++) {
                const double v[3] = {(sb & 1 ? -1 : 1) * rep[t][0], (sb & 2 ? -1 : 1) * rep[t][1],
                                     (sb & 4 ? -1 : 1) * rep[t][2]};
                int bi = 0;
                double bd = -1e30;
                for (int i = 0; i < 20; i++) {
                    const double d = v[0] * vm[i][0] + v[1] * vm[i][1] + v[2] * vm[i][2];
                    if (d > bd) bd = d, bi = i;
                }
                lut[t * 8 + sb] = bi;
            }
        S3D_CUDA(e, cudaMemcpyToSymbol(c_face_lut, lut, sizeof(lut)));
    }
    if (!e->d_mesh) S3D_CUDA(e, cudaMalloc(&e->d_mesh, sizeof(MeshDev)));
    S3D_CUDA(e, cudaMemcpyAsync(e->d_mesh, &M, sizeof(M), cudaMemcpyHostToDevice, e->stream));
    S3D_CUDA(e, cudaStreamSynchronize(e->stream));
    e->have_mesh = true;
    return 0;
}

int s3d_k_orient_list(s3d_engine *e, s3d_keypoint *d_kp, int n, double sig_fctr,
                      double corner_thresh, unsigned char *d_ok, double *d_conf)
{
    if (n <= 0) return 0;
    const PyrTable T = make_table(e);
    k_orient<false><<<(n + 127) / 128, 128, 0, e->stream>>>(d_kp, n, T, sig_fctr, corner_thresh,
                                                           d_ok, d_conf, nullptr, nullptr);
    S3D_LAUNCH_CHECK(e);
    return 0;
}

// Window-weight tables for the keypoint levels s = 0..K-1 of every octave (see k_orient_table).
static int build_orient_tables(s3d_engine *e, const PyrTable &T, double sig_fctr)
{
    const int L = e->noct * e->nlev_g;
    {   // tables and lists depend on the levels' scales and units only: keep them between calls
        std::vector<double> key;
        key.push_back(sig_fctr);
        key.push_back((double)e->K);
        key.push_back((double)e->first_level);
        for (int lv = 0; lv < L; lv++) {
            const s3d_geom &g = e->slab.empty() ? e->g[lv].g : e->slab_g[lv];
            key.push_back(g.scale), key.push_back(g.ux), key.push_back(g.uy), key.push_back(g.uz);
        }
        if (e->d_ori_tabs && key == e->ori_key) return e->ori_key_rc;
        e->ori_key = key;
        e->ori_key_rc = 1;
    }
    std::vector<OriTab> tabs(L);
    size_t total = 0;
    e->ori_max_twx = 0;
    e->ori_lists_ok = true;
    for (int lv = 0; lv < L; lv++) {
        OriTab &t = tabs[lv];
        t.off = -1;
        t.rx = t.ry = t.rz = 0;
        t.list_n = 0;
        t.pad[0] = t.pad[1] = t.pad[2] = 0;
        const int sidx = lv % e->nlev_g + e->first_level;  // level index s
        if (sidx < 0 || sidx >= e->K) continue;
        const s3d_geom &g = e->slab.empty() ? e->g[lv].g : e->slab_g[lv];
        const double rad = sig_fctr * g.scale * 3.0;
        const double r[3] = {rad / (double)(float)g.ux, rad / (double)(float)g.uy,
                             rad / (double)(float)g.uz};
        if (!(r[0] < 200 && r[1] < 200 && r[2] < 200)) return 1;  // absurd scales: literal path
        t.rx = (int)ceil(r[0]) + 1;
        t.ry = (int)ceil(r[1]) + 1;
        t.rz = (int)ceil(r[2]) + 1;
        if (t.rx > 127 || t.ry > 127 || t.rz > 127) e->ori_lists_ok = false;  // offsets are signed bytes
        e->ori_max_twx = std::max(e->ori_max_twx, 2 * t.rx + 1);
        const size_t n = (size_t)(2 * t.rx + 1) * (2 * t.ry + 1) * (2 * t.rz + 1);
        if (total + n > ((size_t)1 << 27)) return 1;
        t.off = (int)total;
        total += n;
    }
    if (total == 0) return 1;
    if (total > e->ori_pool_cap) {
        if (e->d_ori_pool) cudaFree(e->d_ori_pool);
        if (e->d_ori_lists) cudaFree(e->d_ori_lists);
        e->d_ori_pool = nullptr;
        e->d_ori_lists = nullptr;
        e->ori_pool_cap = 0;
        S3D_CUDA(e, cudaMalloc(&e->d_ori_pool, total * sizeof(float)));
        S3D_CUDA(e, cudaMalloc(&e->d_ori_lists, total * sizeof(int2)));
        e->ori_pool_cap = total;
    }
    if ((size_t)L > e->ori_tabs_cap) {
        if (e->d_ori_tabs) cudaFree(e->d_ori_tabs);
        e->d_ori_tabs = nullptr;
        e->ori_tabs_cap = 0;
        S3D_CUDA(e, cudaMalloc(&e->d_ori_tabs, L * sizeof(OriTab)));
        e->ori_tabs_cap = L;
    }
    S3D_CUDA(e, cudaMemcpyAsync(e->d_ori_tabs, tabs.data(), L * sizeof(OriTab), cudaMemcpyHostToDevice,
                                e->stream));
    S3D_CUDA(e, cudaStreamSynchronize(e->stream));  // `tabs` is a stack-owned host buffer
    k_orient_table<<<dim3(64, L), 256, 0, e->stream>>>(static_cast<const OriTab *>(e->d_ori_tabs), T,
                                                      sig_fctr, e->d_ori_pool);
    S3D_LAUNCH_CHECK(e);
    if (e->ori_lists_ok) {
        k_orient_list<<<L, 1024, 0, e->stream>>>(static_cast<OriTab *>(e->d_ori_tabs), e->d_ori_pool,
                                                 static_cast<int2 *>(e->d_ori_lists));
        S3D_LAUNCH_CHECK(e);
    }
    e->ori_key_rc = 0;
    return 0;
}

int s3d_pack_candidates(s3d_engine *e, s3d_keypoint *d_out)
{
    const int n = e->ncand;
    if (n <= 0) return 0;
    const PyrTable T = make_table(e);
    k_cand_to_keypoints<<<(n + 255) / 256, 256, 0, e->stream>>>(e->d_cand, n, T.scales, e->nlev_g,
                                                               e->first_level, d_out);
    S3D_LAUNCH_CHECK(e);
    return 0;
}

int s3d_k_orientations(s3d_engine *e, double corner_thresh)
{
    const int n = e->ncand;
    if (n <= 0) {
        e->nkp = 0;
        return 0;
    }
    if (s3d_pack_candidates(e, e->d_kp_all)) return -1;
    // ori_sig_fctr = 1.5 (sift.c:51, 1281)
    {
        if (s3d_gradients_prepare(e)) return -1;
        const PyrTable T = make_table(e);
        const int trc = build_orient_tables(e, T, 1.5);
        if (trc < 0) return -1;
        if (trc == 0) {  // every keypoint level has a table
            bool grads = e->grad_valid;
            for (int lv = 0; grads && lv < (int)e->g.size(); lv++) {
                const int sidx = lv % e->nlev_g + e->first_level;
                if (sidx >= 0 && sidx < e->K && !e->grad[lv]) grads = false;
            }
            if (e->ori_lists_ok && !e->opt_orient_v1) {  // grouped kernel
                const int G = e->opt_orient_g == 16 || e->opt_orient_g == 32 ? e->opt_orient_g : 8;
                const int per_cta = ORI_WARPS * (32 / G);
                const int grid = (n + per_cta - 1) / per_cta;
                const bool sc = !grads || e->opt_orient_scalar;
                const OriTab *tb = static_cast<const OriTab *>(e->d_ori_tabs);
                const int2 *ls = static_cast<const int2 *>(e->d_ori_lists);
#define S3D_ORI_LAUNCH(GG, SC) \
    k_orient_group<GG, SC><<<grid, 32 * ORI_WARPS, 0, e->stream>>>(e->d_kp_all, n, T, 1.5, corner_thresh, e->d_ok, tb, ls)
                if (G == 8 && !sc) S3D_ORI_LAUNCH(8, false);
                else if (G == 8) S3D_ORI_LAUNCH(8, true);
                else if (G == 16 && !sc) S3D_ORI_LAUNCH(16, false);
                else if (G == 16) S3D_ORI_LAUNCH(16, true);
                else if (!sc) S3D_ORI_LAUNCH(32, false);
                else S3D_ORI_LAUNCH(32, true);
#undef S3D_ORI_LAUNCH
            } else if (e->opt_orient_batch == 8)
                k_orient<true, 8><<<(n + 127) / 128, 128, 0, e->stream>>>(
                    e->d_kp_all, n, T, 1.5, corner_thresh, e->d_ok, nullptr,
                    static_cast<const OriTab *>(e->d_ori_tabs), e->d_ori_pool);
            else
                k_orient<true, 4><<<(n + 127) / 128, 128, 0, e->stream>>>(
                    e->d_kp_all, n, T, 1.5, corner_thresh, e->d_ok, nullptr,
                    static_cast<const OriTab *>(e->d_ori_tabs), e->d_ori_pool);
        } else {
            k_orient<false><<<(n + 127) / 128, 128, 0, e->stream>>>(
                e->d_kp_all, n, T, 1.5, corner_thresh, e->d_ok, nullptr, nullptr, nullptr);
        }
        S3D_LAUNCH_CHECK(e);
    }
    k_flag_scan<<<1, 1024, 0, e->stream>>>(e->d_ok, n, e->d_pos, e->d_counter);
    S3D_LAUNCH_CHECK(e);
    k_compact_keypoints<<<(n + 255) / 256, 256, 0, e->stream>>>(e->d_kp_all, e->d_ok, e->d_pos, n,
                                                               e->d_kp);
    S3D_LAUNCH_CHECK(e);
    return 0;
}

int s3d_k_descriptors(s3d_engine *e, const s3d_keypoint *d_kp, int n, unsigned char *d_out,
                      int kp_levels_only)
{
    if (n <= 0) return 0;
    if (!e->have_mesh) return s3d_fail(e, "mesh not set", cudaSuccess, __FILE__, __LINE__);
    if (s3d_gradients_prepare(e)) return -1;
    const PyrTable T = make_table(e);
    // k_descriptor3 needs the gradient volume (and its block maxima) of every level a keypoint
    // can name: the caller vouches that all keypoints sit on keypoint levels s = 0..K-1, and all
    // of those must have got their volume (s3d_gradients_prepare is best effort)
    bool v3 = kp_levels_only && e->grad_valid && !e->opt_desc_v1 && !e->opt_desc_path && !e->opt_desc_v2;
    for (int lv = 0; v3 && lv < (int)e->g.size(); lv++) {
        const int sidx = lv % e->nlev_g + e->first_level;
        if (sidx >= 0 && sidx < e->K && !e->grad[lv]) v3 = false;
    }
    if (v3) {
        if (!(e->opt_icos_fast & 1))
            k_descriptor3<4, false, false><<<n, D3_THREADS, 0, e->stream>>>(d_kp, n, T, e->d_mesh, d_out);
        else if (e->opt_desc_occ == 3)
            k_descriptor3<3, true, true><<<n, D3_THREADS, 0, e->stream>>>(d_kp, n, T, e->d_mesh, d_out);
        else if (e->opt_desc_pre)
            k_descriptor3<4, true, true><<<n, D3_THREADS, 0, e->stream>>>(d_kp, n, T, e->d_mesh, d_out);
        else
            k_descriptor3<4, true, false><<<n, D3_THREADS, 0, e->stream>>>(d_kp, n, T, e->d_mesh, d_out);
        S3D_LAUNCH_CHECK(e);
        return 0;
    }
    // v2 packs window offsets in 10 bits per axis; a window wider than 1023 voxels (absurd
    // scales) takes the simple kernel
    if (e->opt_desc_v1)
        k_descriptor<<<n, DESC_THREADS, 0, e->stream>>>(d_kp, n, T, e->d_mesh, d_out,
                                                        e->opt_icos_fast);
    else if (e->opt_desc_path)
        k_descriptor2<4, true><<<n, DESC2_THREADS, 0, e->stream>>>(
            d_kp, n, T, e->d_mesh, d_out, (e->opt_icos_fast & 1) | (e->opt_desc_path << 1));
    else if (e->opt_desc_occ == 3)
        k_descriptor2<3, false><<<n, DESC2_THREADS, 0, e->stream>>>(d_kp, n, T, e->d_mesh, d_out,
                                                                   e->opt_icos_fast & 1);
    else
        k_descriptor2<4, false><<<n, DESC2_THREADS, 0, e->stream>>>(
            d_kp, n, T, e->d_mesh, d_out, (e->opt_icos_fast & 1) | (e->opt_desc_norot ? 16 : 0));
    S3D_LAUNCH_CHECK(e);
    return 0;
}

int s3d_k_dense(s3d_engine *e, const float *d_smooth, const float *, int nx, int ny, int nz,
                const float inv_units[3], float *d_temp12)
{
    if (!e->have_mesh) return s3d_fail(e, "mesh not set", cudaSuccess, __FILE__, __LINE__);
    const size_t total = (size_t)nx * ny * nz;
    const size_t want = (total + 255) / 256;
    const size_t cap = (size_t)e->num_sms * 16;
    k_dense_bary<<<(int)(want < cap ? want : cap), 256, 0, e->stream>>>(
        d_smooth, nx, ny, nz, inv_units[0], inv_units[1], inv_units[2], e->d_mesh, d_temp12);
    S3D_LAUNCH_CHECK(e);
    return 0;
}

int s3d_k_dense_rotate(s3d_engine *e, const float *d_smooth, int nx, int ny, int nz,
                       const float units[3], double ori_sigma, double desc_sigma,
                       double corner_thresh, float *d_out12)
{
    const size_t total = (size_t)nx * ny * nz;
    const size_t blocks = (total + DROT_THREADS - 1) / DROT_THREADS;
    if (blocks == 0 || blocks > 0x7fffffffull)
        return s3d_fail(e, "dense rotate: volume size", cudaSuccess, __FILE__, __LINE__);
    k_dense_rotate<<<(unsigned)blocks, DROT_THREADS, 0, e->stream>>>(
        d_smooth, nx, ny, nz, units[0], units[1], units[2], ori_sigma, desc_sigma, corner_thresh,
        e->d_mesh, d_out12);
    S3D_LAUNCH_CHECK(e);
    return 0;
}

int s3d_k_dense_post(s3d_engine *e, float *d_desc12, const float *d_raw, size_t nvox)
{
    const size_t want = (nvox + 255) / 256;
    const size_t cap = (size_t)e->num_sms * 16;
    k_dense_post<<<(int)(want < cap ? want : cap), 256, 0, e->stream>>>(d_desc12, d_raw, nvox);
    S3D_LAUNCH_CHECK(e);
    return 0;
}
