/* csv_format_check.c -- sift3d_b200_format_f (host/csv_io.c) against snprintf("%f"): random bit
 * patterns of every exponent the fast path takes, every tie pattern m / 2^k (k <= 12) whose
 * seventh decimal is an exact 5, values one ulp either side of (j + 0.5) * 1e-6, floats widened
 * to double (what a SIFT3D_FLOAT matrix passes), denormals, signed zeros, NaN / Inf, huge values.
 *   gcc -O2 -I include -I sift3d_b200/host tools/csv_format_check.c sift3d_b200/host/csv_io.c -lz -lm -fopenmp
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int sift3d_b200_format_f(char *dst, double v);

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint64_t rnd(void)
{
    uint64_t z = (rng_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static long bad = 0, total = 0;
static void check(double v)
{
    char a[512], b[512];
    const int n = sift3d_b200_format_f(a, v);
    const int m = snprintf(b, sizeof(b), "%f", v);
    total++;
    if (n != m || memcmp(a, b, (size_t)n)) {
        if (bad++ < 10) {
            a[n < 500 ? n : 500] = 0;
            printf("MISMATCH %a: got %s want %s\n", v, a, b);
        }
    }
}

int main(int argc, char **argv)
{
    const long nrand = argc > 1 ? atol(argv[1]) : 4000000;
    long i;
    int k;
    for (i = 0; i < nrand; i++) { /* random mantissa and sign, exponent in [-40, 70] around 1.0 */
        uint64_t bits = rnd();
        const uint64_t ex = 1023 - 40 + rnd() % 111;
        double v;
        bits = (bits & 0x800fffffffffffffull) | (ex << 52);
        memcpy(&v, &bits, 8);
        check(v);
    }
    for (i = 0; i < nrand / 4; i++) { /* any bit pattern at all */
        uint64_t bits = rnd();
        double v;
        memcpy(&v, &bits, 8);
        check(v);
    }
    for (i = 0; i < nrand / 4; i++) { /* floats in [0, 1): descriptor values */
        const float f = (float)(rnd() >> 40) * (1.0f / 16777216.0f) * (i & 1 ? 1.0f : 0.03333f);
        check((double)f);
        check(-(double)f);
    }
    for (k = 0; k <= 12; k++) /* exact ties and near-ties */
        for (i = -5000; i <= 5000; i++) {
            const double v = ldexp((double)i, -k);
            check(v);
            check(nextafter(v, 1e9));
            check(nextafter(v, -1e9));
        }
    for (i = 0; i < 200000; i++) {
        const double v = ((double)(rnd() % 2000000) + 0.5) * 1e-6;
        check(v);
        check(nextafter(v, 1e9));
        check(nextafter(v, -1e9));
        check(-v);
    }
    {
        const double sp[] = {0.0, -0.0, 4.9e-324, -4.9e-324, 2.2250738585072014e-308, 5e-7, -5e-7, 4.999999e-7,
                             5.000001e-7, 1e-6, 0.9999995, 0.99999949999, 999999.9999995, 8796093022207.9, 8796093022208.0,
                             -8796093022208.5, 9.2e18, 1e22, 1.7976931348623157e308, -1.7976931348623157e308,
                             INFINITY, -INFINITY, NAN, -NAN, 123456789.1234565, 0.0078125, 0.0234375, 2.5e-6, 3.5e-6};
        for (i = 0; i < (long)(sizeof(sp) / sizeof(sp[0])); i++) check(sp[i]);
    }
    printf("csv_format_check: %ld values, %ld mismatches\n", total, bad);
    return bad != 0;
}
