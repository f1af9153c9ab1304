// Host-side check of the branch-free noise core (float64 selection, FP32 contributions) against the
// float64 oracle (no GPU needed): with the reference's own decisions there are no candidate-set flips,
// so the error is FP32 rounding of the contributions only, at every frequency, exact ties included.
//   g++ -O2 -mfma -ffp-contract=off -I. tools/host_noise_check.cpp oracle/libnixis_oracle.so -Wl,-rpath,$PWD/oracle -o /tmp/host_noise_check
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <random>
#include "../nixis_b200/csrc/nxb_noise3_fast.cuh"
extern "C" {
void nxo_init(int64_t seed, int32_t *perm, int32_t *pgi);
double nxo_noise3(double x, double y, double z, const int32_t *perm, const int32_t *pgi);
}
int main(int argc, char **argv)
{
    int64_t seed = argc > 1 ? atoll(argv[1]) : 12345;
    int32_t perm[256], pgi[256];
    nxo_init(seed, perm, pgi);
    uint8_t perm8[256], grad8[256];
    for (int i = 0; i < 256; ++i) {
        perm8[i] = (uint8_t)perm[i];
        int g = pgi[i] / 3, q = g / 3, a = g % 3;
        grad8[i] = (uint8_t)(((q & 1) ? 0 : 1) | ((q & 2) ? 2 : 0) | ((q & 4) ? 4 : 0) | (a << 3));
    }
    std::vector<char> sm(NXF_SMEM_BYTES);
    nxf_build_tables(perm8, grad8, sm.data(), 0, 1);
    {   // branch-free selection == literal selection, exactly, including ties
        std::mt19937_64 r2(11);
        std::uniform_real_distribution<float> U01(0.0f, 1.0f);
        long n = 0, bad = 0;
        auto check = [&](float fx, float fy, float fz) {
            float fsum = fx + fy + fz;
            float A[8], B[8]; int a0, a1, b0, b1;
            nxf_select_branchy(fx, fy, fz, fsum, A[0], A[1], A[2], A[3], A[4], A[5], A[6], A[7], a0, a1);
            nxf_select(fx, fy, fz, fsum, B[0], B[1], B[2], B[3], B[4], B[5], B[6], B[7], b0, b1);
            bool ok = a0 == b0 && a1 == b1;
            for (int i = 0; i < 8; ++i) ok = ok && A[i] == B[i];
            ++n; if (!ok) { if (bad < 5) printf("  MISMATCH f=(%g %g %g) e=(%d,%d) vs (%d,%d)\n", fx, fy, fz, a0, a1, b0, b1); ++bad; }
        };
        for (int i = 0; i < 4000000; ++i) check(U01(r2), U01(r2), U01(r2));
        const float grid[] = {0.0f, 0.125f, 0.25f, 1.0f / 3.0f, 0.375f, 0.5f, 0.625f, 2.0f / 3.0f, 0.75f, 0.875f, 0.99999994f};
        for (float x : grid) for (float y : grid) for (float z : grid) check(x, y, z);
        printf("selection: %ld cases, %ld mismatches\n", n, bad);
    }
    std::mt19937_64 rng(7);
    std::uniform_real_distribution<double> U(-1, 1);
    const double scales[] = {1.5, 9.4, 58.6, 366.2, 915.5, 5722.0, 35763.0};
    for (double f : scales) {
        double maxerr = 0, sumerr = 0; int nbig = 0; const int N = 400000;
        double worst[3] = {0, 0, 0};
        for (int i = 0; i < N; ++i) {
            double x = U(rng), y = U(rng), z = U(rng);
            double n = std::sqrt(x * x + y * y + z * z); x /= n; y /= n; z /= n;
            if (i % 7 == 0) { y = x; n = std::sqrt(x * x + y * y + z * z); x /= n; y /= n; z /= n; }   // tie-prone: symmetric
            if (i % 11 == 0) { z = 0; n = std::sqrt(x * x + y * y); x /= n; y /= n; }
            uint32_t lane4 = (uint32_t)(i & 31) * 4;
            // float64 coordinates in, as the fBm kernel passes them (terrain.py:17 verts * n_roughness)
            float got = nxf_noise3_x103_d(x * f, y * f, z * f, sm.data(), lane4) * (1.0f / 103.0f);
            double ref = nxo_noise3(x * f, y * f, z * f, perm, pgi);
            double e = std::fabs((double)got - ref);
            sumerr += e;
            if (e > maxerr) { maxerr = e; worst[0] = x * f; worst[1] = y * f; worst[2] = z * f; }
            if (e > 2e-6) ++nbig;
        }
        printf("f=%9.1f  max %.3e  mean %.3e  n(>2e-6) %d  worst at (%.4f %.4f %.4f)\n", f, maxerr, sumerr / N, nbig, worst[0], worst[1], worst[2]);
    }
    // lattice-aligned and tie-prone points
    double maxerr = 0; int n = 0, nflip = 0;
    for (int a = -8; a <= 8; ++a) for (int b = -8; b <= 8; ++b) for (int c = -8; c <= 8; ++c) {
        double x = a * 0.25 + (a % 3 ? 0.0 : 0.001), y = b * 0.25, z = c * 0.25 - (c % 2 ? 0.002 : 0.0);
        float got = nxf_noise3_x103_d(x, y, z, sm.data(), (uint32_t)(n & 31) * 4) * (1.0f / 103.0f);
        double ref = nxo_noise3(x, y, z, perm, pgi);
        double e = std::fabs(got - ref);
        if (e > maxerr) maxerr = e;
        if (e > 2e-6) ++nflip;
        ++n;
    }
    printf("quarter-lattice points: n=%d max %.3e n(>2e-6) %d\n", n, maxerr, nflip);
    return 0;
}
