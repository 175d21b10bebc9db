// Stand-alone check and timing of the int8 error-free contraction (autogp.jl_b200/csrc/agp_ozaki.cu) on one B200.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/ozaki_update_test tools/ozaki_update_test.cu
//   tools/ozaki_update_test            (GPU box)
//
// (1) row scales and digit planes against the host;  (2) T_ik -= sum_{j<c0} L_ij L_kj^T against a long-double host
// reference on a small batch;  (3) time of every launch of a W-wide super-column schedule at n = 2048 x 64.
#include "../autogp.jl_b200/csrc/agp_ozaki.cu"

#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define CK(x)                                                                               \
    do {                                                                                    \
        cudaError_t e_ = (x);                                                               \
        if (e_ != cudaSuccess) {                                                            \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                                        \
        }                                                                                   \
    } while (0)

static double urand() { return (double)rand() / RAND_MAX * 2.0 - 1.0; }

int main(int argc, char** argv) {
    CK(cudaSetDevice(0));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    CK(agp::configure_ozaki());
    int* d_err;
    CK(cudaMalloc(&d_err, 4));
    CK(cudaMemset(d_err, 0, 4));

    {
        // ---- (1), (2): small batch ------------------------------------------------------------------
        const int P = 3, nt = 5, ld = nt * 128, c0 = 2, c1 = 4;
        const long long ms = (long long)ld * ld;
        std::vector<double> L((size_t)P * ms, 0.0), Kd((size_t)P * ld);
        srand(11);
        for (int p = 0; p < P; ++p)
            for (int r = 0; r < ld; ++r) {
                double s2 = 0;
                for (int c = 0; c < r; ++c) {
                    // wide dynamic range within a row, as in a real factor (entries decay away from the diagonal)
                    const double x = urand() * exp(-0.02 * (r - c) * (1 + p)) * (0.3 + p);
                    L[(size_t)p * ms + (size_t)r * ld + c] = x;
                    s2 += x * x;
                }
                Kd[(size_t)p * ld + r] = s2 + 0.01 + fabs(urand());
                L[(size_t)p * ms + (size_t)r * ld + r] = Kd[(size_t)p * ld + r];  // the Gram diagonal, as after the fill
            }
        double *dL, *dR;
        int8_t* dS;
        CK(cudaMalloc(&dL, L.size() * 8));
        CK(cudaMalloc(&dR, (size_t)P * ld * 16));
        CK(cudaMalloc(&dS, (size_t)agp::OZ_SLICES * P * ms));
        CK(cudaMemset(dS, 0, (size_t)agp::OZ_SLICES * P * ms));
        CK(cudaMemcpy(dL, L.data(), L.size() * 8, cudaMemcpyHostToDevice));
        agp::launch_ozaki_rowscale(dL, ms, ld, P, dR, 0);
        std::vector<double> R((size_t)P * ld * 2);
        CK(cudaMemcpy(R.data(), dR, R.size() * 8, cudaMemcpyDeviceToHost));
        long long bad_scale = 0;
        for (int w = 0; w < P * ld; ++w) {
            const double sc = R[2 * w];
            if (!(0.99 * sc >= sqrt(Kd[w]) * 0.9999999 && sc < 2.05 * sqrt(Kd[w])) || R[2 * w + 1] * sc != 0x1p55) ++bad_scale;
        }
        printf("(1a) row scales: %lld of %d outside [sqrt K_rr / 0.99, 2.05 sqrt K_rr)\n", bad_scale, P * ld);
        // the trailing tiles now hold a running Schur complement T (any values): random
        std::vector<double> T = L;
        for (int p = 0; p < P; ++p)
            for (int r = c0 * 128; r < ld; ++r)
                for (int c = c0 * 128; c < ld; ++c) T[(size_t)p * ms + (size_t)r * ld + c] = urand();
        CK(cudaMemcpy(dL, T.data(), T.size() * 8, cudaMemcpyHostToDevice));
        agp::launch_ozaki_slice(dL, ms, ld, nt, P, dR, dS, 0, c0, c0, 0);
        std::vector<int8_t> S((size_t)agp::OZ_SLICES * P * ms);
        CK(cudaMemcpy(S.data(), dS, S.size(), cudaMemcpyDeviceToHost));
        long long bad_digit = 0;
        double worst_rep = 0;
        for (int p = 0; p < P; ++p)
            for (int r = c0 * 128; r < ld; ++r)
                for (int c = 0; c < c0 * 128; ++c) {
                    const double x = T[(size_t)p * ms + (size_t)r * ld + c], sc = R[2 * ((size_t)p * ld + r)];
                    long double rep = 0;
                    for (int q = 0; q < agp::OZ_SLICES; ++q) {
                        const int d = S[(size_t)q * P * ms + (size_t)p * ms + (size_t)r * ld + c];
                        if (q == 0 && (d < -127 || d > 127)) ++bad_digit;
                        rep += (long double)d * ldexpl(1.0L, -7 - 8 * q);
                    }
                    const double err = fabs((double)(rep * (long double)sc - (long double)x)) / sc;
                    if (err > worst_rep) worst_rep = err;
                }
        printf("(1b) digit planes: %lld leading digits outside [-127, 127]; worst |sum_p a_p 2^(-7-8p) - x / 2^e| = %.3e (bound 2^-56 = %.3e)\n", bad_digit,
               worst_rep, ldexp(1.0, -56));

        agp::OzakiMaps maps;
        if (!agp::make_ozaki_maps(dS, ld, P, &maps)) {
            printf("tensor map encode failed\n");
            return 1;
        }
        for (int variant : {2, 3}) {
        CK(cudaMemcpy(dL, T.data(), T.size() * 8, cudaMemcpyHostToDevice));
        agp::OzakiParams prm{dL, ms, ld, nt, P, dR, c0, c1, d_err, 2000000000ull};
        agp::launch_ozaki_update(prm, maps, sms, 0, variant);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        int err = 0;
        CK(cudaMemcpy(&err, d_err, 4, cudaMemcpyDeviceToHost));
        std::vector<double> out(T.size());
        CK(cudaMemcpy(out.data(), dL, out.size() * 8, cudaMemcpyDeviceToHost));
        double worst = 0, worst_fp64 = 0, maxres = 0;
        long long touched_wrong = 0;
        for (int p = 0; p < P; ++p)
            for (int r = c0 * 128; r < ld; ++r)
                for (int c = c0 * 128; c < ld; ++c) {
                    const size_t o = (size_t)p * ms + (size_t)r * ld + c;
                    const bool in = c < c1 * 128 && c <= r;  // lower tiles of block columns [c0, c1); diagonal tiles: lower triangle only
                    if (!in) {
                        if (out[o] != T[o]) ++touched_wrong;
                        continue;
                    }
                    long double s = 0;
                    double s64 = 0;
                    for (int j = 0; j < c0 * 128; ++j) {
                        const double a = T[(size_t)p * ms + (size_t)r * ld + j], b = T[(size_t)p * ms + (size_t)c * ld + j];
                        s += (long double)a * (long double)b;
                        s64 = fma(a, b, s64);
                    }
                    const long double ref = (long double)T[o] - s;
                    maxres = fmax(maxres, fabs((double)ref));
                    worst = fmax(worst, fabs((double)((long double)out[o] - ref)));
                    worst_fp64 = fmax(worst_fp64, fabs((double)((long double)(T[o] - s64) - ref)));
                }
        printf("(2) variant %d: update, P = %d, nt = %d, block columns [%d, %d), depth %d: err flag %d, max |T| = %.3f, max error int8 path = %.3e, "
               "FP64 fma chain = %.3e; %lld entries outside the target region changed\n",
               variant, P, nt, c0, c1, c0 * 128, err, maxres, worst, worst_fp64, touched_wrong);
        }
        cudaFree(dL);
        cudaFree(dR);
        cudaFree(dS);
    }

    {
        // ---- (3): timing at the benchmark size ------------------------------------------------------
        const int P = (argc > 1) ? atoi(argv[1]) : 64, nt = (argc > 2) ? atoi(argv[2]) : 16, ld = nt * 128;
        const long long ms = (long long)ld * ld;
        double *dL, *dR;
        int8_t* dS;
        CK(cudaMalloc(&dL, (size_t)P * ms * 8));
        CK(cudaMalloc(&dR, (size_t)P * ld * 16));
        CK(cudaMalloc(&dS, (size_t)agp::OZ_SLICES * P * ms));
        CK(cudaMemset(dS, 1, (size_t)agp::OZ_SLICES * P * ms));
        CK(cudaMemset(dL, 0, (size_t)P * ms * 8));
        std::vector<double> R((size_t)P * ld * 2);
        for (size_t w = 0; w < (size_t)P * ld; ++w) R[2 * w] = 1.0, R[2 * w + 1] = 0x1p55;
        CK(cudaMemcpy(dR, R.data(), R.size() * 8, cudaMemcpyHostToDevice));
        agp::OzakiMaps maps;
        if (!agp::make_ozaki_maps(dS, ld, P, &maps)) {
            printf("tensor map encode failed\n");
            return 1;
        }
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        for (int variant : {2, 3})
        for (int W : {4}) {
            double total_ms = 0, total_slice = 0, total_flop = 0;
            for (int c0 = W; c0 < nt; c0 += W) {
                const int c1 = (c0 + W < nt) ? c0 + W : nt;
                agp::OzakiParams prm{dL, ms, ld, nt, P, dR, c0, c1, d_err, 2000000000ull};
                agp::launch_ozaki_update(prm, maps, sms, 0, variant);  // warm
                CK(cudaEventRecord(e0));
                agp::launch_ozaki_update(prm, maps, sms, 0, variant);
                CK(cudaEventRecord(e1));
                CK(cudaDeviceSynchronize());
                float t = 0;
                CK(cudaEventElapsedTime(&t, e0, e1));
                CK(cudaEventRecord(e0));
                agp::launch_ozaki_slice(dL, ms, ld, nt, P, dR, dS, c0 - W, c0, c0, 0);
                CK(cudaEventRecord(e1));
                CK(cudaDeviceSynchronize());
                float ts = 0;
                CK(cudaEventElapsedTime(&ts, e0, e1));
                long long items = 0;
                for (int k = c0; k < c1; ++k) items += 2 * (nt - k);
                items *= P;
                const double flop = 2.0 * items * 128.0 * 64.0 * (c0 * 128.0);  // FP64-equivalent
                printf("    W = %d, block columns [%2d, %2d): %6lld items, depth %4d: update %.3f ms = %.1f FP64-equivalent TFLOP/s (%.0f int8 TOP/s), slicing the previous super-column %.3f ms\n",
                       W, c0, c1, items, c0 * 128, t, flop / (t * 1e-3) * 1e-12, 28.0 * flop / (t * 1e-3) * 1e-12, ts);
                total_ms += t;
                total_slice += ts;
                total_flop += flop;
            }
            printf("(3) variant %d: P = %d, n = %d, W = %d: all updates %.3f ms (%.1f FP64-equivalent TFLOP/s), all slicing %.3f ms\n", variant, P, ld, W, total_ms,
                   total_flop / (total_ms * 1e-3) * 1e-12, total_slice);
        }
#if OZ_STATS
        for (int variant : {2, 3})
        for (int c0 : {4, 8, 12}) {
            long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0}, st[8];
            CK(cudaMemcpyToSymbol(agp::oz_stats, z, sizeof(z)));
            agp::OzakiParams prm{dL, ms, ld, nt, P, dR, c0, c0 + 4, d_err, 2000000000ull};
            agp::launch_ozaki_update(prm, maps, sms, 0, variant);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpyFromSymbol(st, agp::oz_stats, sizeof(st)));
            const double it = (double)st[5];
            printf("    stats variant %d c0 = %d: per unit (clocks): MMA thread total %.0f, wait B %.0f, wait A %.0f, wait TMEM free %.0f; epilogue (all passes) %.0f; ideal tensor time %d\n", variant, c0,
                   st[0] / it, st[1] / it, st[2] / it, st[3] / it, st[4] / it, c0 * 112 * 64);
        }
#endif
        int err = 0;
        CK(cudaMemcpy(&err, d_err, 4, cudaMemcpyDeviceToHost));
        printf("    err flag %d\n", err);
    }
    return 0;
}
