// match.cu -- brute-force descriptor matching for sm_100a.
//
// Replaces SIFT3D_nn_match / match_desc (sift.c:2840-2969): for every descriptor of one
// store, the nearest and second-nearest descriptor of the other store by SSD over the 768
// histogram values, the ratio test ssd_best / ssd_nearest > nn_thresh^2, and the
// forward-backward consistency check.
//
// Exactness.  The reference accumulates the SSD in f64, sequentially over the 768 values,
// diff = (double)a - (double)b; ssd += diff * diff (no FMA on baseline x86-64).  This kernel
// keeps exactly that: one f64 accumulator per (i, j) pair, k ascending, separately rounded
// multiply and add -- so every SSD is bit-identical to the CPU's, and with them argmin (first
// index on ties, strict `<`, sift.c:2946), the second-smallest value and the ratio decision.
// The reference's early exit (sift.c:2941) only skips work: a partial sum that already exceeds
// ssd_nearest can change neither the best nor the nearest.
// A ||a||^2 + ||b||^2 - 2ab GEMM on the tensor cores would be ~30x faster but cannot reproduce
// near-tie decisions; B200's 64 f64 lanes/clk/SM make the exact form affordable
// (3 f64 ops per value and pair).
//
// Layout: a CTA owns 64 rows of A and sweeps all of B in 64-column tiles; 16x16 threads, 4x4
// pairs each (rows ty*4.., columns tx, tx+16, ..); tiles are staged in shared memory already
// widened to f64, [k][record]: row reads are warp broadcasts, column reads consecutive words.  Per row the running
// (best, index, second) is merged across the 16 threads of the row with shuffles.
#include "common.cuh"

#include <cfloat>

namespace {

constexpr int TM = 64, TN = 64, KC = 32, NT = 256;
constexpr int NUMEL = S3D_DESC_NUMEL;
constexpr int STRIDE_F = S3D_DESC_STRIDE / 4;  // floats per descriptor record

struct Best {
    double best, second;
    int idx;
};

__device__ __forceinline__ void best_init(Best &b)
{
    b.best = DBL_MAX;
    b.second = DBL_MAX;
    b.idx = -1;
}

// candidate (ssd, j) in ascending j order (sift.c:2946-2952)
__device__ __forceinline__ void best_push(Best &b, double ssd, int j)
{
    if (ssd < b.best) {
        b.second = b.best;
        b.best = ssd;
        b.idx = j;
    } else {
        b.second = fmin(b.second, ssd);
    }
}

// merge two partial results over disjoint index sets (ties keep the smaller index, which is
// what the sequential scan with its strict `<` does)
__device__ __forceinline__ Best best_merge(const Best &lo, const Best &hi)
{
    Best r;
    if (hi.best < lo.best || (hi.best == lo.best && (unsigned)hi.idx < (unsigned)lo.idx)) {
        r.best = hi.best;
        r.idx = hi.idx;
        r.second = fmin(lo.best, hi.second);
    } else {
        r.best = lo.best;
        r.idx = lo.idx;
        r.second = fmin(lo.second, hi.best);
    }
    return r;
}

// A: nA records, B: nB records (S3D_DESC_STRIDE bytes each, 768 floats first).
// out_idx[i] = index of the nearest B for A[i], or -1 if the ratio test rejects it
// (match_desc, sift.c:2893-2969).
__global__ void __launch_bounds__(NT) k_nn_pass(const float *__restrict__ A, int nA,
                                                const float *__restrict__ B, int nB,
                                                float thresh2, int *__restrict__ out_idx)
{
    // +1: the transposing tile store (32 consecutive k of one record per warp) would otherwise
    // hit one bank 32 times
    __shared__ double As[KC][TM + 1];
    __shared__ double Bs[KC][TN + 1];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int row0 = blockIdx.x * TM;
    Best run[4];
#pragma unroll
    for (int r = 0; r < 4; r++) best_init(run[r]);

    for (int col0 = 0; col0 < nB; col0 += TN) {
        double acc[4][4];
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) acc[r][c] = 0.0;
        for (int k0 = 0; k0 < NUMEL; k0 += KC) {
            __syncthreads();
            // 64 records x 32 floats per tile: thread -> (record = e / 32, k = e % 32)
            for (int e = tid; e < TM * KC; e += NT) {
                const int rec = e >> 5, k = e & 31;
                const int ia = row0 + rec, ib = col0 + rec;
                As[k][rec] = ia < nA ? (double)__ldg(A + (size_t)ia * STRIDE_F + k0 + k) : 0.0;
                Bs[k][rec] = ib < nB ? (double)__ldg(B + (size_t)ib * STRIDE_F + k0 + k) : 0.0;
            }
            __syncthreads();
#pragma unroll 4
            for (int k = 0; k < KC; k++) {
                double a[4], b[4];
#pragma unroll
                for (int r = 0; r < 4; r++) a[r] = As[k][ty * 4 + r];
#pragma unroll
                for (int c = 0; c < 4; c++) b[c] = Bs[k][tx + 16 * c];  // lanes: consecutive words
#pragma unroll
                for (int r = 0; r < 4; r++)
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        const double d = __dsub_rn(a[r], b[c]);
                        acc[r][c] = __dadd_rn(acc[r][c], __dmul_rn(d, d));
                    }
            }
        }
        // this tile's candidates: own 4 columns (tx, tx+16, ..: ascending), then across tx
#pragma unroll
        for (int r = 0; r < 4; r++) {
            Best t;
            best_init(t);
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int j = col0 + tx + 16 * c;
                if (j < nB) best_push(t, acc[r][c], j);
            }
            // tree over the 16 threads of the row (lanes tx = 0..15 of a half warp)
#pragma unroll
            for (int o = 1; o < 16; o <<= 1) {
                Best p;
                p.best = __shfl_down_sync(0xffffffffu, t.best, o, 16);
                p.second = __shfl_down_sync(0xffffffffu, t.second, o, 16);
                p.idx = __shfl_down_sync(0xffffffffu, t.idx, o, 16);
                if ((tx & (2 * o - 1)) == 0) t = best_merge(t, p);
            }
            if (tx == 0) run[r] = best_merge(run[r], t);
        }
    }
    if (tx == 0) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int i = row0 + ty * 4 + r;
            if (i >= nA) continue;
            // sift.c:2955: reject when the nearest neighbour is too close (f64 division; a NaN
            // from 0/0 compares false, i.e. accepts, like the CPU)
            const bool reject = __ddiv_rn(run[r].best, run[r].second) > (double)thresh2;
            out_idx[i] = reject ? -1 : run[r].idx;
        }
    }
}

// forward-backward consistency (sift.c:2876-2884)
__global__ void k_nn_consistency(const int *__restrict__ fwd, int n1, const int *__restrict__ bwd,
                                 int *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n1) return;
    const int m = fwd[i];
    out[i] = (m >= 0 && bwd[m] == i) ? m : -1;
}

}  // namespace

extern "C" {

int s3d_nn_match_device(s3d_engine *e, const void *dev_d1, int n1, const void *dev_d2, int n2,
                        float nn_thresh, int *dev_matches)
{
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(e->device);
    int rc = -1;
    int *tmp = nullptr;
    do {
        if (n1 < 1 || n2 < 1) {
            s3d_fail(e, "s3d_nn_match: empty descriptor store", cudaSuccess, __FILE__, __LINE__);
            break;
        }
        if (cudaMalloc(&tmp, ((size_t)n1 + n2) * sizeof(int)) != cudaSuccess) {
            s3d_fail(e, "s3d_nn_match: cudaMalloc", cudaGetLastError(), __FILE__, __LINE__);
            break;
        }
        const float t2 = nn_thresh * nn_thresh;  // f32 product, as in sift.c:2955
        const float *A = static_cast<const float *>(dev_d1), *B = static_cast<const float *>(dev_d2);
        k_nn_pass<<<(n1 + TM - 1) / TM, NT, 0, e->stream>>>(A, n1, B, n2, t2, tmp);
        e->launches++;
        k_nn_pass<<<(n2 + TM - 1) / TM, NT, 0, e->stream>>>(B, n2, A, n1, t2, tmp + n1);
        e->launches++;
        k_nn_consistency<<<(n1 + 255) / 256, 256, 0, e->stream>>>(tmp, n1, tmp + n1, dev_matches);
        e->launches++;
        cudaError_t ce = cudaGetLastError();
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
        if (ce != cudaSuccess) {
            s3d_fail(e, "s3d_nn_match: kernels", ce, __FILE__, __LINE__);
            break;
        }
        rc = 0;
    } while (0);
    if (tmp) cudaFree(tmp);
    if (prev >= 0) cudaSetDevice(prev);
    return rc;
}

int s3d_nn_match(s3d_engine *e, const void *host_d1, int n1, const void *host_d2, int n2,
                 float nn_thresh, int *host_matches)
{
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(e->device);
    unsigned char *d1 = nullptr, *d2 = nullptr;
    int *dm = nullptr;
    int rc = -1;
    do {
        if (n1 < 1 || n2 < 1) {
            s3d_fail(e, "s3d_nn_match: empty descriptor store", cudaSuccess, __FILE__, __LINE__);
            break;
        }
        const size_t b1 = (size_t)n1 * S3D_DESC_STRIDE, b2 = (size_t)n2 * S3D_DESC_STRIDE;
        if (cudaMalloc(&d1, b1) != cudaSuccess || cudaMalloc(&d2, b2) != cudaSuccess ||
            cudaMalloc(&dm, (size_t)n1 * sizeof(int)) != cudaSuccess) {
            s3d_fail(e, "s3d_nn_match: cudaMalloc", cudaGetLastError(), __FILE__, __LINE__);
            break;
        }
        if (cudaMemcpyAsync(d1, host_d1, b1, cudaMemcpyHostToDevice, e->stream) != cudaSuccess ||
            cudaMemcpyAsync(d2, host_d2, b2, cudaMemcpyHostToDevice, e->stream) != cudaSuccess) {
            s3d_fail(e, "s3d_nn_match: upload", cudaGetLastError(), __FILE__, __LINE__);
            break;
        }
        if (s3d_nn_match_device(e, d1, n1, d2, n2, nn_thresh, dm)) break;
        if (cudaMemcpyAsync(host_matches, dm, (size_t)n1 * sizeof(int), cudaMemcpyDeviceToHost,
                            e->stream) != cudaSuccess ||
            cudaStreamSynchronize(e->stream) != cudaSuccess) {
            s3d_fail(e, "s3d_nn_match: download", cudaGetLastError(), __FILE__, __LINE__);
            break;
        }
        rc = 0;
    } while (0);
    cudaStreamSynchronize(e->stream);
    if (d1) cudaFree(d1);
    if (d2) cudaFree(d2);
    if (dm) cudaFree(dm);
    if (prev >= 0) cudaSetDevice(prev);
    return rc;
}

}  // extern "C"
