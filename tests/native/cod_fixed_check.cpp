// CPU check of romis_b200/csrc/cod_fixed.hpp against include/romis_cod.h: same bits for every system (tests/test_cod_fixed.py).
//   g++ -O2 -ffp-contract=off -std=c++17 -Iinclude -Iromis_b200/csrc tests/native/cod_fixed_check.cpp -o check && ./check
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>
extern "C" {
#include "romis_cod.h"
}
#include "cod_fixed.hpp"

static uint32_t bits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }

template <int N> static long check(uint64_t seed, int cases, long& by_rank_total, long* by_rank) {
    std::mt19937_64 rng(seed);
    std::uniform_real_distribution<float> U(0.0f, 1.0f);
    long bad = 0;
    for (int t = 0; t < cases; t++) {
        float A[N * N] = {0};
        const int kind = t % 12;
        // technique matrices are sums of outer products v v^T of non-negative vectors (render.cpp:209-214)
        int terms = kind < 4 ? 2 * N : (kind < 9 ? 1 + (int)(rng() % N) : 0);
        for (int s = 0; s < terms; s++) {
            float v[N];
            for (int i = 0; i < N; i++) v[i] = (rng() % 4 == 0) ? 0.0f : U(rng) * (kind == 3 ? 1e-18f : 1.0f);
            if (kind == 2 && s > 0) for (int i = 0; i < N; i++) v[i] *= 1e-4f;            // ill-conditioned
            if (kind == 5) v[N - 1] = v[0];                                             // repeated technique: equal rows / columns
            for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) A[i * N + j] += v[i] * v[j];
        }
        if (kind == 9) for (int i = 0; i < N; i++) A[i * N + i] = (float)(i % 2);      // diagonal with zeros
        if (kind == 10) { for (int i = 0; i < N * N; i++) A[i] = U(rng) - 0.5f; }        // not symmetric at all
        if (kind == 11) { for (int i = 0; i < N * N; i++) A[i] = U(rng); A[(rng() % N) * N + rng() % N] = (t % 24 == 11) ? NAN : INFINITY; }
        float B[3][N];
        for (int ch = 0; ch < 3; ch++) for (int i = 0; i < N; i++) B[ch][i] = (rng() % 5 == 0) ? 0.0f : U(rng) * 3.0f;

        romis_cod ref; romis_cod_compute(&ref, A, N);
        float xr[3][N];
        for (int ch = 0; ch < 3; ch++) romis_cod_solve(&ref, B[ch], xr[ch]);

        romis::CodFixed<N> d;
        for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) d.qr[j][i] = A[i * N + j];
        float xf[3][N];
        romis::cod_fixed_solve3<N>(d, [&](int ch, float* b) { for (int i = 0; i < N; i++) b[i] = B[ch][i]; },
                                   [&](int ch, const float* x) { for (int i = 0; i < N; i++) xf[ch][i] = x[i]; });
        bool ok = d.rank == ref.rank;
        for (int ch = 0; ch < 3 && ok; ch++) for (int i = 0; i < N; i++) if (bits(xr[ch][i]) != bits(xf[ch][i]) && !(xr[ch][i] != xr[ch][i] && xf[ch][i] != xf[ch][i])) ok = false;   // any NaN = any NaN: x86 keeps an operand's sign / payload, the GPU returns the canonical one
        if (!ok) { if (bad < 5) std::printf("N=%d case %d kind %d: rank %d vs %d, x0 %a vs %a\n", N, t, kind, ref.rank, d.rank, xr[0][0], xf[0][0]); bad++; }
        by_rank[ref.rank]++; by_rank_total++;
    }
    return bad;
}

int main(int argc, char** argv) {
    const int cases = argc > 1 ? std::atoi(argv[1]) : 200000;
    long total = 0, by_rank[12] = {0};
    long bad = check<6>(1, cases, total, by_rank) + check<2>(2, cases / 4, total, by_rank) + check<4>(3, cases / 4, total, by_rank) + check<11>(4, cases / 20, total, by_rank);
    std::printf("systems %ld, by rank:", total);
    for (int r = 0; r < 12; r++) std::printf(" %d:%ld", r, by_rank[r]);
    std::printf("\nmismatches %ld\n", bad);
    return bad != 0;
}
