// Host check of the window enumeration of k_descriptor3 (sift3d_b200/csrc/desc_cell_geom.cuh).
//
// For random keypoints (rotation, scale, sub-voxel centre, anisotropic units, windows clipped by
// the volume border, near-axis-aligned and exactly axis-aligned rotations) the set of voxels
// reached through  cells x bounding rows x scanned intervals x d3_member  must equal the set the
// reference visits (sphere test + 0 <= vb < 4, sift.c:1866-1881), each voxel exactly once.
// Also reports how tight the scan is (voxels scanned / voxels accepted, rows scanned / rows hit).
//
//   nvcc -O2 -o desc_cell_host_check tools/desc_cell_host_check.cu && ./desc_cell_host_check
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../sift3d_b200/csrc/desc_cell_geom.cuh"

static void sphere_bounds(float c, float rad, float uf, int n, int &lo, int &hi)
{
    volatile float q = rad / uf;
    volatile float a = c - q, b = c + q;
    lo = (int)fmaxf(floorf(a), 1.0f);
    hi = (int)fminf(ceilf(b), (float)(n - 2));
}

static void random_rotation(std::mt19937 &g, int mode, float R[9])
{
    std::normal_distribution<double> N(0, 1);
    double q[4];
    double n = 0;
    for (int i = 0; i < 4; i++) q[i] = N(g), n += q[i] * q[i];
    n = sqrt(n);
    for (int i = 0; i < 4; i++) q[i] /= n;
    if (mode == 1) q[0] = 1, q[1] = q[2] = q[3] = 0;               // identity
    if (mode == 2) q[0] = 1, q[1] = 1e-4, q[2] = -2e-4, q[3] = 0;  // nearly axis aligned
    if (mode == 3) q[0] = q[1] = sqrt(0.5), q[2] = q[3] = 0;       // 90 degrees about x
    n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    const double w = q[0] / n, x = q[1] / n, y = q[2] / n, z = q[3] / n;
    const double M[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                         2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                         2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)};
    for (int i = 0; i < 9; i++) R[i] = (float)M[i];
}

int main(int argc, char **argv)
{
    const int ncase = argc > 1 ? atoi(argv[1]) : 400;
    std::mt19937 g(12345);
    std::uniform_real_distribution<double> U(0, 1);
    long long tot_ref = 0, tot_scan = 0, tot_rows = 0, tot_rows_hit = 0, bad_cases = 0;
    for (int c = 0; c < ncase; c++) {
        const int nx = 96, ny = 90, nz = 84;
        D3Key K;
        const int mode = c % 8 == 5 ? 1 : (c % 8 == 6 ? 2 : (c % 8 == 7 ? 3 : 0));
        float R[9];
        random_rotation(g, mode, R);
        if (c % 50 == 49) R[0] *= 1.05f;  // not orthonormal: whole-window fallback
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) K.Rt[3 * i + j] = R[3 * j + i];
        const bool integer_centre = c % 3 != 0;
        K.kx = (float)(4 + U(g) * (nx - 8)), K.ky = (float)(4 + U(g) * (ny - 8)), K.kz = (float)(4 + U(g) * (nz - 8));
        if (integer_centre) K.kx = floorf(K.kx), K.ky = floorf(K.ky), K.kz = floorf(K.kz);
        const bool aniso = c % 5 == 4;
        K.ux = aniso ? 0.7f : 1.0f, K.uy = aniso ? 1.3f : 1.0f, K.uz = aniso ? 0.9f : 1.0f;
        const double sd = 1.6 * pow(2.0, (c % 3) / 3.0) * (c % 7 == 6 ? 0.5 : 1.0);
        const float sigma = (float)(sd * 7.071067812);
        const float win_radius = (float)(2.0 * (double)sigma);
        K.half = (float)((double)win_radius / sqrt(2.0));
        const float desc_width = D3_MUL(2.0f, K.half);
        K.hw = desc_width / 4.0f;
        K.binf = 1.0f / K.hw;
        K.r2 = D3_MUL(win_radius, win_radius);
        sphere_bounds(K.kx, win_radius, K.ux, nx, K.x0, K.x1);
        sphere_bounds(K.ky, win_radius, K.uy, ny, K.y0, K.y1);
        sphere_bounds(K.kz, win_radius, K.uz, nz, K.z0, K.z1);
        d3_key_finish(K);
        D3Scan Q;
        d3_scan_setup(K, Q);

        std::vector<unsigned char> ref((size_t)nx * ny * nz, 0), got((size_t)nx * ny * nz, 0);
        // the reference's visit set (any cell)
        long long nref = 0;
        for (int z = K.z0; z <= K.z1; z++)
            for (int y = K.y0; y <= K.y1; y++)
                for (int x = K.x0; x <= K.x1; x++) {
                    const float vx = D3_MUL(D3_SUB((float)x, K.kx), K.ux), vy = D3_MUL(D3_SUB((float)y, K.ky), K.uy),
                                vz = D3_MUL(D3_SUB((float)z, K.kz), K.uz);
                    const float sq = D3_ADD(D3_ADD(D3_MUL(vx, vx), D3_MUL(vy, vy)), D3_MUL(vz, vz));
                    if (sq > K.r2) continue;
                    bool in = true;
                    for (int a = 0; a < 3; a++) {
                        const float vk = D3_ADD(D3_ADD(D3_MUL(K.Rt[3 * a], vx), D3_MUL(K.Rt[3 * a + 1], vy)),
                                                D3_MUL(K.Rt[3 * a + 2], vz));
                        const float vb = D3_MUL(D3_ADD(vk, K.half), K.binf);
                        in = in && !(vb < 0.0f || vb >= 4.0f);
                    }
                    if (in) ref[x + (size_t)nx * (y + (size_t)ny * z)] = 1, nref++;
                }
        long long nscan = 0, nacc = 0, dup = 0;
        for (int cell = 0; cell < 64; cell++) {
            D3Cell C;
            d3_cell_bbox(K, cell & 3, (cell >> 2) & 3, cell >> 4, C);
            for (int z = C.zlo; z <= C.zhi; z++)
                for (int y = C.ylo; y <= C.yhi; y++) {
                    int xa, cnt;
                    d3_scan_row(Q, C.ibf, y, z, xa, cnt);
                    tot_rows++;
                    bool hit = false;
                    for (int x = xa; x < xa + cnt; x++) {
                        float sq, dv[3];
                        nscan++;
                        if (!d3_member(K, C.ibf, K.r2, (float)x, (float)y, (float)z, sq, dv)) continue;
                        unsigned char &gg = got[x + (size_t)nx * (y + (size_t)ny * z)];
                        if (gg) dup++;
                        gg = 1;
                        nacc++;
                        hit = true;
                    }
                    tot_rows_hit += hit;
                }
        }
        const bool same = memcmp(ref.data(), got.data(), ref.size()) == 0;
        if (!same || dup || nacc != nref) {
            long long miss = 0, extra = 0;
            for (size_t i = 0; i < ref.size(); i++) miss += ref[i] && !got[i], extra += got[i] && !ref[i];
            printf("case %d (mode %d aniso %d int %d): MISMATCH ref %lld got %lld dup %lld missing %lld extra %lld\n",
                   c, mode, (int)aniso, (int)integer_centre, nref, nacc, dup, miss, extra);
            bad_cases++;
        }
        tot_ref += nref;
        tot_scan += nscan;
    }
    printf("%d cases, %lld voxels visited by the reference, %lld scanned (x%.4f), rows scanned %lld hit %lld (%.1f%%)\n",
           ncase, tot_ref, tot_scan, (double)tot_scan / (double)tot_ref, tot_rows, tot_rows_hit,
           100.0 * (double)tot_rows_hit / (double)tot_rows);
    printf(bad_cases ? "FAILED: %lld cases differ\n" : "all cases identical (%lld bad)\n", bad_cases);
    return bad_cases ? 1 : 0;
}
