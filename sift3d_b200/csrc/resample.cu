// resample.cu -- image resampling under an affine transform.
//
// Replaces im_inv_transform (imutil.c:2040-2081) for an Affine tform (apply_Affine_xyz,
// imutil.c:2651-2672) with resample_linear (imutil.c:2085-2124, trilinear, zero outside
// [0, n-1]) or resample_lanczos2 (imutil.c:2127-2178), i.e. what im_resample
// (imutil.c:2191-2244) and register_SIFT3D_resample / `--warped` run.
//
// One thread per output voxel, all arithmetic in f64 in the reference's expression order with
// separately rounded multiplies and adds (baseline x86-64 has no FMA), so LINEAR is
// bit-identical to the CPU.  LANCZOS2 evaluates sin() with CUDA's libm instead of glibc's
// (both faithfully rounded, not identical): agreement to ~1e-15 relative before the f32 store.
#include "common.cuh"

#include <algorithm>
#include <cfloat>

namespace {

struct Affine34 {
    double a[12];  // row-major 3 x 4
};

__device__ __forceinline__ double dm(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double da(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double ds(double a, double b) { return __dsub_rn(a, b); }

__device__ __forceinline__ double lanczos2(double x)
{  // lanczos(x, a = 2), imutil.c:2181-2185
    const double pi_x = dm(M_PI, x);
    return __ddiv_rn(dm(dm(2.0, sin(pi_x)), sin(__ddiv_rn(pi_x, 2.0))), dm(pi_x, pi_x));
}

template <int INTERP>
__global__ void __launch_bounds__(256) k_resample(const float *__restrict__ src, int nx, int ny,
                                                  int nz, int nc, const Affine34 T,
                                                  float *__restrict__ dst, int dnx, int dny,
                                                  int dnz)
{
    const size_t total = (size_t)dnx * dny * dnz;
    const size_t sxs = nc, sys = (size_t)nc * nx, szs = (size_t)nc * nx * ny;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int xi = (int)(idx % dnx);
        const size_t r = idx / dnx;
        const int yi = (int)(r % dny), zi = (int)(r / dny);
        const double xd = xi, yd = yi, zd = zi;
        // apply_Affine_xyz: ((a0*x + a1*y) + a2*z) + a3
        const double x = da(da(da(dm(T.a[0], xd), dm(T.a[1], yd)), dm(T.a[2], zd)), T.a[3]);
        const double y = da(da(da(dm(T.a[4], xd), dm(T.a[5], yd)), dm(T.a[6], zd)), T.a[7]);
        const double z = da(da(da(dm(T.a[8], xd), dm(T.a[9], yd)), dm(T.a[10], zd)), T.a[11]);
        float *out = dst + idx * nc;
        const bool oob = x < 0 || x > nx - 1 || y < 0 || y > ny - 1 || z < 0 || z > nz - 1;
        if (oob) {  // also catches NaN coordinates?  no: NaN compares false, like the CPU
            for (int c = 0; c < nc; c++) out[c] = 0.0f;
            continue;
        }
        if (INTERP == 0) {
            const int fx = (int)floor(x), fy = (int)floor(y), fz = (int)floor(z);
            const int cx = (int)ceil(x), cy = (int)ceil(y), cz = (int)ceil(z);
            const double dx = ds(x, (double)fx), dy = ds(y, (double)fy), dz = ds(z, (double)fz);
            const double ox = ds(1.0, dx), oy = ds(1.0, dy), oz = ds(1.0, dz);
            for (int c = 0; c < nc; c++) {
                const float *p = src + c;
                const double c0 = __ldg(p + fx * sxs + fy * sys + fz * szs);
                const double c1 = __ldg(p + fx * sxs + cy * sys + fz * szs);
                const double c2 = __ldg(p + cx * sxs + fy * sys + fz * szs);
                const double c3 = __ldg(p + cx * sxs + cy * sys + fz * szs);
                const double c4 = __ldg(p + fx * sxs + fy * sys + cz * szs);
                const double c5 = __ldg(p + fx * sxs + cy * sys + cz * szs);
                const double c6 = __ldg(p + cx * sxs + fy * sys + cz * szs);
                const double c7 = __ldg(p + cx * sxs + cy * sys + cz * szs);
                double o = dm(dm(dm(c0, ox), oy), oz);
                o = da(o, dm(dm(dm(c1, ox), dy), oz));
                o = da(o, dm(dm(dm(c2, dx), oy), oz));
                o = da(o, dm(dm(dm(c3, dx), dy), oz));
                o = da(o, dm(dm(dm(c4, ox), oy), dz));
                o = da(o, dm(dm(dm(c5, ox), dy), dz));
                o = da(o, dm(dm(dm(c6, dx), oy), dz));
                o = da(o, dm(dm(dm(c7, dx), dy), dz));
                out[c] = (float)o;
            }
        } else {
            const double a = 2.0;
            const int x0 = (int)fmax(floor(x) - a, 0.0), x1 = (int)fmin(floor(x) + a, (double)(nx - 1));
            const int y0 = (int)fmax(floor(y) - a, 0.0), y1 = (int)fmin(floor(y) + a, (double)(ny - 1));
            const int z0 = (int)fmax(floor(z) - a, 0.0), z1 = (int)fmin(floor(z) + a, (double)(nz - 1));
            for (int c = 0; c < nc; c++) {
                double val = 0.0;
                for (int zs = z0; zs <= z1; zs++)
                    for (int ys = y0; ys <= y1; ys++)
                        for (int xs = x0; xs <= x1; xs++) {
                            const double xw = da(fabs(ds((double)xs, x)), DBL_EPSILON);
                            const double yw = da(fabs(ds((double)ys, y)), DBL_EPSILON);
                            const double zw = da(fabs(ds((double)zs, z)), DBL_EPSILON);
                            const double k = dm(dm(lanczos2(xw), lanczos2(yw)), lanczos2(zw));
                            val = da(val, dm(k, (double)__ldg(src + c + xs * sxs + ys * sys + zs * szs)));
                        }
                out[c] = (float)val;
            }
        }
    }
}

}  // namespace

extern "C" {

int s3d_resample_affine_device(s3d_engine *e, const float *dev_src, int nx, int ny, int nz, int nc,
                               const double A[12], int interp, float *dev_dst, int dnx, int dny,
                               int dnz)
{
    if (nx < 1 || ny < 1 || nz < 1 || nc < 1 || dnx < 1 || dny < 1 || dnz < 1 ||
        (interp != 0 && interp != 1))
        return s3d_fail(e, "s3d_resample_affine: bad arguments (interp: 0 LINEAR, 1 LANCZOS2)",
                        cudaSuccess, __FILE__, __LINE__);
    Affine34 T;
    for (int i = 0; i < 12; i++) T.a[i] = A[i];
    const size_t total = (size_t)dnx * dny * dnz;
    const size_t want = (total + 255) / 256;
    const int grid = (int)std::min<size_t>(want, (size_t)e->num_sms * 16);
    if (interp == 0)
        k_resample<0><<<grid, 256, 0, e->stream>>>(dev_src, nx, ny, nz, nc, T, dev_dst, dnx, dny, dnz);
    else
        k_resample<1><<<grid, 256, 0, e->stream>>>(dev_src, nx, ny, nz, nc, T, dev_dst, dnx, dny, dnz);
    S3D_LAUNCH_CHECK(e);
    return 0;
}

int s3d_resample_affine(s3d_engine *e, const float *host_src, int nx, int ny, int nz, int nc,
                        const double A[12], int interp, float *host_dst, int dnx, int dny, int dnz)
{  // contiguous channel-interleaved host volumes (im_default_stride, imutil.c:1453)
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(e->device);
    const size_t nb_src = (size_t)nx * ny * nz * nc * sizeof(float);
    const size_t nb_dst = (size_t)dnx * dny * dnz * nc * sizeof(float);
    float *ds = nullptr, *dd = nullptr;
    int rc = -1;
    do {
        if (cudaMalloc(&ds, nb_src) != cudaSuccess || cudaMalloc(&dd, nb_dst) != cudaSuccess) {
            s3d_fail(e, "s3d_resample_affine: cudaMalloc", cudaGetLastError(), __FILE__, __LINE__);
            break;
        }
        if (cudaMemcpyAsync(ds, host_src, nb_src, cudaMemcpyHostToDevice, e->stream) != cudaSuccess) {
            s3d_fail(e, "s3d_resample_affine: upload", cudaGetLastError(), __FILE__, __LINE__);
            break;
        }
        if (s3d_resample_affine_device(e, ds, nx, ny, nz, nc, A, interp, dd, dnx, dny, dnz)) break;
        if (cudaMemcpyAsync(host_dst, dd, nb_dst, cudaMemcpyDeviceToHost, e->stream) != cudaSuccess ||
            cudaStreamSynchronize(e->stream) != cudaSuccess) {
            s3d_fail(e, "s3d_resample_affine: download", cudaGetLastError(), __FILE__, __LINE__);
            break;
        }
        rc = 0;
    } while (0);
    cudaStreamSynchronize(e->stream);
    if (ds) cudaFree(ds);
    if (dd) cudaFree(dd);
    if (prev >= 0) cudaSetDevice(prev);
    return rc;
}

}  // extern "C"
