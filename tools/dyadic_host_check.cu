// CPU check of the register-blocked dyadic FIR (sift3d_b200/csrc/blur_dyadic.cu): the same
// __host__ __device__ line code, driven by host loops, against the oracle's apply_Sep_FIR_filter
// restatement (oracle/sift3d_oracle.c).  TEST TOOL (loads the oracle).
//   nvcc -O2 -I include tools/dyadic_host_check.cu oracle/_build/liboracle.so -o /tmp/dyc && /tmp/dyc
#include "../sift3d_b200/csrc/blur_dyadic.cu"
#include "../oracle/sift3d_oracle.h"

#include <cstdlib>
#include <cstring>
#include <vector>

int s3d_fail(s3d_engine *, const char *, cudaError_t, const char *, int) { return -1; }

template <int AXIS, int O, int HW, int RUN>
static void host_axis(const float *src, float *dst, int nx, int ny, int nz, const TapSet &taps)
{
    constexpr int P = 1 << O;
    constexpr int UHW = (HW + P - 1) >> O;
    const int n = AXIS == 0 ? nx : (AXIS == 1 ? ny : nz);
    const int nruns = (n + RUN - 1) / RUN;
    const size_t st = AXIS == 0 ? 1 : (AXIS == 1 ? (size_t)nx : (size_t)nx * ny);
    const size_t nthreads = (size_t)nx * ny * nz / n * nruns;
    const int start = UHW, end = n - 1 - (UHW + 1);
    const float uf = 1.0f / (float)P;
    for (size_t t = 0; t < nthreads; t++) {
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
        if (i0 < start || i0 + RUN - 1 > end)
            for (int k = 0; k < RUN; k++) {
                const int i = i0 + k;
                if (i < n && (i < start || i > end)) acc[k] = boundary_point(line, st, n, i, taps, uf);
            }
        float *o = dst + line_off + (size_t)i0 * st;
        for (int k = 0; k < RUN; k++)
            if (i0 + k < n) o[(size_t)k * st] = acc[k];
    }
}

template <int O, int HW>
static void host_xtile(const float *src, float *dst, int nx, int ny, int nz, const TapSet &taps)
{
    using T = XTile<O, HW>;
    const size_t nrows = (size_t)ny * nz;
    const int ntx = (nx + XT_OUT - 1) / XT_OUT;
    const size_t nblocks = (size_t)ntx * ((nrows + XT_ROWS - 1) / XT_ROWS);
    std::vector<float> s(XT_ROWS * T::WP);
    std::vector<float> accs(XT_OUT * XT_RUN);
    for (size_t b = 0; b < nblocks; b++) {
        const int x_base = (int)(b % ntx) * XT_OUT;
        const size_t row0 = (b / ntx) * XT_ROWS;
        for (auto &v : s) v = -777.0f;
        for (int tid = 0; tid < XT_OUT; tid++) xtile_load<O, HW>(tid, s.data(), src, nx, nrows, x_base, row0);
        for (int tid = 0; tid < XT_OUT; tid++) {
            float acc[XT_RUN] = {0};
            xtile_compute<O, HW>(tid, s.data(), src, nx, nrows, x_base, row0, taps, acc);
            memcpy(&accs[tid * XT_RUN], acc, sizeof(acc));
        }
        for (int tid = 0; tid < XT_OUT; tid++) {
            float acc[XT_RUN];
            memcpy(acc, &accs[tid * XT_RUN], sizeof(acc));
            xtile_stage<O, HW>(tid, s.data(), acc);
        }
        for (int tid = 0; tid < XT_OUT; tid++) xtile_fix_ends<O, HW>(tid, s.data(), src, nx, nrows, x_base, row0, taps);
        for (int tid = 0; tid < XT_OUT; tid++) xtile_store<O, HW>(tid, s.data(), dst, nx, nrows, x_base, row0);
    }
}

template <int O, int HW>
static int check(int nx, int ny, int nz, unsigned seed)
{
    const size_t n = (size_t)nx * ny * nz;
    std::vector<float> src(n), a(n), b(n), c(n), ref(n);
    srand(seed);
    for (size_t i = 0; i < n; i++) src[i] = (float)rand() / (float)RAND_MAX - 0.3f;
    TapSet taps;
    memset(&taps, 0, sizeof(taps));
    const double sigma = HW / 3.0 - 1e-3;
    taps.width = orc_gauss_taps(sigma, taps.t, S3D_MAX_TAPS);
    if (taps.width != 2 * HW + 1) { printf("width %d != %d\n", taps.width, 2 * HW + 1); return 1; }
    const double u = (double)(1 << O);
    const double units[3] = {u, u, u};
    orc_blur(src.data(), ref.data(), nx, ny, nz, 1, units, taps.t, taps.width, 1.0);
    host_xtile<O, HW>(src.data(), a.data(), nx, ny, nz, taps);
    host_axis<1, O, HW, 16>(a.data(), b.data(), nx, ny, nz, taps);
    host_axis<2, O, HW, 16>(b.data(), c.data(), nx, ny, nz, taps);
    size_t bad = 0;
    for (size_t i = 0; i < n; i++) bad += memcmp(&c[i], &ref[i], 4) != 0;
    printf("O=%d HW=%d %dx%dx%d: %zu / %zu differ\n", O, HW, nx, ny, nz, bad, n);
    return bad != 0;
}

// z pass over a plane range of a slab buffer vs the same planes of the whole-volume z pass
template <int O, int HW>
static int check_zr(int nx, int ny, int nz, unsigned seed)
{
    constexpr int P = 1 << O, H = ((HW + P - 1) >> O) + 1;
    const size_t plane = (size_t)nx * ny, n = plane * nz;
    std::vector<float> src(n), whole(n);
    srand(seed);
    for (size_t i = 0; i < n; i++) src[i] = (float)rand() / (float)RAND_MAX - 0.3f;
    TapSet taps;
    memset(&taps, 0, sizeof(taps));
    taps.width = orc_gauss_taps(HW / 3.0 - 1e-3, taps.t, S3D_MAX_TAPS);
    host_axis<2, O, HW, 16>(src.data(), whole.data(), nx, ny, nz, taps);
    size_t bad = 0, cnt = 0;
    const int ranges[4][2] = {{0, nz / 3}, {nz / 3, 2 * nz / 3 + 1}, {2 * nz / 3 + 1, nz}, {0, nz}};
    for (auto &rg : ranges) {
        const int o0 = rg[0], o1 = rg[1];                       // global output planes
        const int g0 = o0 - H < 0 ? 0 : o0 - H, g1 = o1 + H > nz ? nz : o1 + H;  // buffer = outputs + halo
        const int nbuf = g1 - g0, zb = o0 - g0, ze = o1 - g0;
        std::vector<float> buf(src.begin() + plane * g0, src.begin() + plane * g1), out(plane * nbuf, -555.0f);
        const size_t nruns = (size_t)(ze - zb + 15) / 16;
        for (size_t t = 0; t < plane * nruns; t++)
            zrange_run<O, HW, 16>(buf.data() + t % plane, out.data() + t % plane, plane, nbuf,
                                  zb + (int)(t / plane) * 16, ze, g0, nz, taps);
        for (int z = o0; z < o1; z++)
            for (size_t i = 0; i < plane; i++, cnt++)
                bad += memcmp(&out[(size_t)(z - g0) * plane + i], &whole[(size_t)z * plane + i], 4) != 0;
    }
    printf("zrange O=%d HW=%d %dx%dx%d: %zu / %zu differ\n", O, HW, nx, ny, nz, bad, cnt);
    return bad != 0;
}

// interleaved channels: the x pass is a "y" pass over lines of stride nc, y / z see rows of nx*nc
template <int HW>
static int check_nc(int nx, int ny, int nz, int nc, unsigned seed)
{
    const size_t n = (size_t)nx * ny * nz * nc;
    std::vector<float> src(n), a(n), b(n), c(n), ref(n);
    srand(seed);
    for (size_t i = 0; i < n; i++) src[i] = (float)rand() / (float)RAND_MAX - 0.3f;
    TapSet taps;
    memset(&taps, 0, sizeof(taps));
    taps.width = orc_gauss_taps(HW / 3.0 - 1e-3, taps.t, S3D_MAX_TAPS);
    if (taps.width != 2 * HW + 1) { printf("width %d != %d\n", taps.width, 2 * HW + 1); return 1; }
    const double units[3] = {1.0, 1.0, 1.0};
    orc_blur(src.data(), ref.data(), nx, ny, nz, nc, units, taps.t, taps.width, 1.0);
    host_axis<1, 0, HW, 16>(src.data(), a.data(), nc, nx, ny * nz, taps);
    host_axis<1, 0, HW, 16>(a.data(), b.data(), nx * nc, ny, nz, taps);
    host_axis<2, 0, HW, 16>(b.data(), c.data(), nx * nc, ny, nz, taps);
    size_t bad = 0;
    for (size_t i = 0; i < n; i++) bad += memcmp(&c[i], &ref[i], 4) != 0;
    printf("nc=%d HW=%d %dx%dx%d: %zu / %zu differ\n", nc, HW, nx, ny, nz, bad, n);
    return bad != 0;
}

int main()
{
    int rc = 0;
    rc |= check_zr<1, 8>(9, 7, 41, 51);
    rc |= check_zr<1, 3>(5, 6, 37, 52);
    rc |= check_zr<2, 8>(6, 5, 40, 53);
    rc |= check_zr<2, 5>(4, 9, 23, 54);
    rc |= check_zr<1, 6>(7, 3, 12, 55);
    rc |= check_nc<9>(22, 20, 18, 12, 41);
    rc |= check_nc<7>(9, 21, 19, 12, 42);
    rc |= check_nc<10>(25, 12, 23, 3, 43);
    rc |= check<0, 9>(41, 22, 20, 44);
    rc |= check<0, 7>(300, 9, 17, 45);
    rc |= check<0, 10>(23, 24, 25, 46);
    rc |= check<0, 2>(37, 21, 19, 21);
    rc |= check<0, 3>(13, 21, 19, 22);
    rc |= check<0, 8>(37, 18, 41, 23);
    rc |= check<0, 5>(8, 8, 8, 24);
    rc |= check<1, 3>(37, 21, 19, 1);
    rc |= check<1, 4>(40, 33, 18, 2);
    rc |= check<1, 5>(16, 17, 35, 3);
    rc |= check<1, 6>(45, 16, 23, 4);
    rc |= check<1, 8>(64, 47, 33, 5);
    rc |= check<2, 3>(37, 21, 19, 6);
    rc |= check<2, 4>(32, 33, 18, 7);
    rc |= check<2, 5>(16, 17, 35, 8);
    rc |= check<2, 6>(45, 16, 23, 9);
    rc |= check<2, 8>(64, 47, 33, 10);
    rc |= check<1, 8>(8, 9, 10, 11);
    rc |= check<0, 8>(300, 5, 7, 31);
    rc |= check<1, 6>(513, 3, 9, 32);
    rc |= check<2, 8>(257, 9, 3, 33);
    rc |= check<1, 3>(256, 4, 5, 34);
    rc |= check<2, 8>(8, 9, 10, 12);
    printf(rc ? "FAIL\n" : "all bit-identical\n");
    return rc;
}
