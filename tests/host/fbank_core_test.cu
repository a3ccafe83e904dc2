// Host-side check of the register-level FFT pieces in ml_vae_b200/csrc/fbank_core.cuh:
// runs the same __host__ __device__ code path as logmel_kernel (pass A, pass B, split)
// on the CPU and compares the 201-bin power spectrum with a float64 DFT.
// Built and run by tests/test_fbank_core_host.py (no GPU needed).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../ml_vae_b200/csrc/fbank_core.cuh"
using namespace mlvae;

int main() {
    std::vector<cpx> tw25(25), tw200(200), tw400(201);
    for (int q = 0; q < 5; ++q) for (int r = 0; r < 5; ++r) { double a = -2.0 * M_PI * (q * r) / 25.0; tw25[q * 5 + r] = {(float)cos(a), (float)sin(a)}; }
    for (int n2 = 0; n2 < 8; ++n2) for (int k1 = 0; k1 < 25; ++k1) { double a = -2.0 * M_PI * (n2 * k1) / 200.0; tw200[n2 * 25 + k1] = {(float)cos(a), (float)sin(a)}; }
    for (int k = 0; k <= 200; ++k) { double a = -2.0 * M_PI * k / 400.0; tw400[k] = {(float)cos(a), (float)sin(a)}; }
    double worst = 0.0;
    srand(1234);
    for (int trial = 0; trial < 8; ++trial) {
        std::vector<float> x(400);
        for (auto &v : x) v = (float)((rand() / (double)RAND_MAX - 0.5) * 0.4);
        if (trial == 7) for (auto &v : x) v = 0.f;
        if (trial == 6) for (int n = 0; n < 400; ++n) x[n] = (float)cos(2.0 * M_PI * 37.0 * n / 400.0);
        // pass A
        std::vector<cpx> Y(200);
        for (int n2 = 0; n2 < 8; ++n2) {
            cpx a[25];
            for (int n1 = 0; n1 < 25; ++n1) a[n1] = {x[16 * n1 + 2 * n2], x[16 * n1 + 2 * n2 + 1]};
            dft25(a, tw25.data());
            for (int k1 = 0; k1 < 25; ++k1) Y[n2 * 25 + k1] = (n2 && k1) ? cmul(a[k1], tw200[n2 * 25 + k1]) : a[k1];
        }
        // pass B + split
        std::vector<float> P(201, -1.f);
        for (int j = 0; j < 13; ++j) {
            cpx y[8], y2[8];
            for (int n2 = 0; n2 < 8; ++n2) y[n2] = Y[n2 * 25 + j];
            dft8(y);
            float pk, pm;
            if (j == 0) {
                split_power(y[0], y[0], tw400[0], pk, pm); P[0] = pk; P[200] = pm;
                for (int k2 = 1; k2 < 4; ++k2) { split_power(y[k2], y[8 - k2], tw400[25 * k2], pk, pm); P[25 * k2] = pk; P[200 - 25 * k2] = pm; }
                split_power(y[4], y[4], tw400[100], pk, pm); P[100] = pk;
            } else {
                for (int n2 = 0; n2 < 8; ++n2) y2[n2] = Y[n2 * 25 + (25 - j)];
                dft8(y2);
                for (int k2 = 0; k2 < 8; ++k2) { int k = j + 25 * k2; split_power(y[k2], y2[7 - k2], tw400[k], pk, pm); P[k] = pk; P[200 - k] = pm; }
            }
        }
        double pmax = 0.0; std::vector<double> ref(201);
        for (int k = 0; k <= 200; ++k) {
            double re = 0, im = 0;
            for (int n = 0; n < 400; ++n) { double a = -2.0 * M_PI * k * n / 400.0; re += x[n] * cos(a); im += x[n] * sin(a); }
            ref[k] = re * re + im * im; pmax = fmax(pmax, ref[k]);
        }
        for (int k = 0; k <= 200; ++k) {
            if (P[k] < 0.f) { printf("FAIL bin %d never written\n", k); return 1; }
            double err = fabs(P[k] - ref[k]) / fmax(pmax, 1e-30);
            if (pmax == 0.0) err = fabs(P[k]);
            worst = fmax(worst, err);
        }
    }
    printf("worst |P - ref| / max(ref) = %.3e\n", worst);
    if (worst > 2e-6) { printf("FAIL\n"); return 1; }
    printf("OK\n");
    return 0;
}
