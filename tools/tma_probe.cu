// tma_probe.cu -- which 3-D tiled TMA configurations does this device accept?  One case per process
// (an illegal instruction poisons the context): tma_probe <nx> <ny> <nz> <bw> <bh> <x> <y> <z> <via>
// via: 0 = descriptor as __grid_constant__ parameter, 1 = descriptor in global memory
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
__device__ __forceinline__ unsigned s32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap pm, const CUtensorMap *gm, int via, int x, int y, int z, int n, float *out)
{
    extern __shared__ __align__(128) unsigned char sm[];
    float *buf = reinterpret_cast<float *>(sm);
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(sm + ((n * 4 + 127) & ~127));
    const unsigned b = s32(bar), d = s32(buf);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 32) {
        const CUtensorMap *m = via ? gm : &pm;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(d),
                     "l"(m), "r"(x), "r"(y), "r"(z), "r"(b)
                     : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(b) : "memory");
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = buf[i];
}
typedef CUresult (*Enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                        const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char **argv)
{
    if (argc < 10) return 2;
    int nx = atoi(argv[1]), ny = atoi(argv[2]), nz = atoi(argv[3]), bw = atoi(argv[4]), bh = atoi(argv[5]);
    int x = atoi(argv[6]), y = atoi(argv[7]), z = atoi(argv[8]), via = atoi(argv[9]);
    std::vector<float> h((size_t)nx * ny * nz);
    for (size_t i = 0; i < h.size(); i++) h[i] = (float)(i % 100003);
    float *d, *out;
    cudaMalloc(&d, h.size() * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    const int n = bw * bh;
    cudaMalloc(&out, n * 4);
    void *fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    CUtensorMap m;
    cuuint64_t dims[3] = {(cuuint64_t)nx, (cuuint64_t)ny, (cuuint64_t)nz}, str[2] = {(cuuint64_t)nx * 4, (cuuint64_t)nx * ny * 4};
    cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}, es[3] = {1, 1, 1};
    CUresult r = ((Enc)fp)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    CUtensorMap *gm;
    cudaMalloc(&gm, sizeof(m));
    cudaMemcpy(gm, &m, sizeof(m), cudaMemcpyHostToDevice);
    const int smem = ((n * 4 + 127) & ~127) + 64;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    k<<<1, 64, smem>>>(m, gm, via, x, y, z, n, out);
    cudaError_t ce = cudaDeviceSynchronize();
    if (ce != cudaSuccess) { printf("FAIL %s\n", cudaGetErrorString(ce)); return 1; }
    std::vector<float> o(n);
    cudaMemcpy(o.data(), out, n * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int j = 0; j < bh; j++)
        for (int i = 0; i < bw; i++) {
            const int gx = x + i, gy = y + j;
            float want = 0.f;
            if (gx >= 0 && gx < nx && gy >= 0 && gy < ny && z >= 0 && z < nz) want = h[((size_t)z * ny + gy) * nx + gx];
            if (o[j * bw + i] != want) bad++;
        }
    printf("%s (%d wrong of %d)\n", bad ? "WRONG" : "OK", bad, n);
    return bad != 0;
}
