// cod_fixed.hpp -- include/romis_cod.h for a system whose size N is a compile-time constant: the same operations in the same
// order (so the same bits: tests/test_cod_fixed.py compares the two on the CPU, over full-rank, rank-deficient and degenerate
// systems), with every array index a compile-time constant after unrolling.  On the GPU the whole decomposition -- the N x N
// factor, the Householder coefficients, the permutation, the right-hand side -- then lives in registers; the generic routine keeps
// them in local memory (1 KB of stack per thread, 400 B of it touched at k = 5: the working set of an SM's resident threads is
// larger than its L1, ncu: 81 % hit rate, 1.3 of 4 warp instructions per cycle).
//
// What is dynamic in the algorithm and how it is made static:
//   * the pivot column `big` of step k: the swap runs over all candidate columns j > k under the predicate j == big;
//   * the permutation built from the transpositions: same;
//   * the rank: everything after the QR factorisation -- the reduction [R11 R12] -> [T11 0] and the three solves -- is a
//     template over the rank R, entered through a switch (lanes of different rank serialise; the ranks of neighbouring
//     pixels mostly agree);
//   * the scatter x[perm[i]] = c[i]: N x N predicated moves.
// Reference: solveSystem (src/rendering/render_utils.h:52) = Eigen's completeOrthogonalDecomposition().solve; the algorithm and
// its citations are in include/romis_cod.h.
#pragma once
#include <float.h>
#include <math.h>

#if defined(__CUDACC__)
#define COD_HD __host__ __device__ __forceinline__
#define COD_HDM __host__ __device__ __forceinline__
#define COD_UNROLL _Pragma("unroll")
#else
#define COD_HD static inline
#define COD_HDM inline
#define COD_UNROLL
#endif

namespace romis {

template <int N> struct CodFixed {
    float qr[N][N];             // [column][row]: R / T above the diagonal, Householder essentials below and right
    float hc[N], zc[N];
    int perm[N];
    int rank;
};

// makeHouseholder pieces shared by the column (QR) and row (Z) reflections: from c0 and the squared norm of the tail
COD_HD void cod_householder_coeffs(float c0, float tailSqNorm, bool& degenerate, float& den, float& tau, float& beta) {
    degenerate = tailSqNorm <= FLT_MIN;
    if (degenerate) { tau = 0.0f; beta = c0; den = 1.0f; }
    else {
        float b = sqrtf(c0 * c0 + tailSqNorm);
        if (c0 >= 0.0f) b = -b;
        den = c0 - b;
        tau = (b - c0) / b;
        beta = b;
    }
}

// romis_cod_factor up to and including the rank (the caller has filled d.qr)
template <int N> COD_HD void cod_fixed_qr(CodFixed<N>& d) {
    const float eps = FLT_EPSILON;
    float normsUpdated[N], normsDirect[N];
    int transp[N];
    float maxNorm = 0.0f;
    COD_UNROLL for (int k = 0; k < N; k++) {
        float s = 0.0f;
        COD_UNROLL for (int i = 0; i < N; i++) s += d.qr[k][i] * d.qr[k][i];
        normsDirect[k] = normsUpdated[k] = sqrtf(s);
        if (k == 0 || normsUpdated[k] > maxNorm) maxNorm = normsUpdated[k];
    }
    const float me = maxNorm * eps;
    const float threshold_helper = (me * me) / (float)N;
    const float norm_downdate_threshold = sqrtf(eps);
    int nonzero_pivots = N;
    float maxpivot = 0.0f;
    COD_UNROLL for (int k = 0; k < N; k++) {
        int big = k; float bigNorm = normsUpdated[k];
        COD_UNROLL for (int j = k + 1; j < N; j++) if (normsUpdated[j] > bigNorm) { bigNorm = normsUpdated[j]; big = j; }
        if (nonzero_pivots == N && bigNorm * bigNorm < threshold_helper * (float)(N - k)) nonzero_pivots = k;
        transp[k] = big;
        // (selects, not `if (j == big) swap`: the optimiser folds the branches back into one access indexed by `big`, and with
        // it the whole factor into local memory)
        COD_UNROLL for (int j = k + 1; j < N; j++) {
            const bool m = j == big;
            COD_UNROLL for (int i = 0; i < N; i++) { const float a = d.qr[k][i], b = d.qr[j][i]; d.qr[k][i] = m ? b : a; d.qr[j][i] = m ? a : b; }
            { const float a = normsUpdated[k], b = normsUpdated[j]; normsUpdated[k] = m ? b : a; normsUpdated[j] = m ? a : b; }
            { const float a = normsDirect[k], b = normsDirect[j]; normsDirect[k] = m ? b : a; normsDirect[j] = m ? a : b; }
        }
        float tailSqNorm = 0.0f;
        COD_UNROLL for (int i = k + 1; i < N; i++) tailSqNorm += d.qr[k][i] * d.qr[k][i];
        bool degenerate; float den, tau, beta;
        cod_householder_coeffs(d.qr[k][k], tailSqNorm, degenerate, den, tau, beta);
        COD_UNROLL for (int i = k + 1; i < N; i++) d.qr[k][i] = degenerate ? 0.0f : d.qr[k][i] / den;
        d.hc[k] = tau;
        d.qr[k][k] = beta;
        if (fabsf(beta) > maxpivot) maxpivot = fabsf(beta);
        if (N - k > 1 && tau != 0.0f) {
            COD_UNROLL for (int j = k + 1; j < N; j++) {
                float tmp = 0.0f;
                COD_UNROLL for (int i = k + 1; i < N; i++) tmp += d.qr[k][i] * d.qr[j][i];
                tmp += d.qr[j][k];
                d.qr[j][k] -= tau * tmp;
                COD_UNROLL for (int i = k + 1; i < N; i++) d.qr[j][i] -= (tau * d.qr[k][i]) * tmp;
            }
        }
        COD_UNROLL for (int j = k + 1; j < N; j++) {
            if (normsUpdated[j] != 0.0f) {
                float temp = fabsf(d.qr[j][k]) / normsUpdated[j];
                temp = (1.0f + temp) * (1.0f - temp);
                temp = temp < 0.0f ? 0.0f : temp;
                const float ratio = normsUpdated[j] / normsDirect[j];
                const float temp2 = temp * (ratio * ratio);
                if (temp2 <= norm_downdate_threshold) {
                    float s = 0.0f;
                    COD_UNROLL for (int i = k + 1; i < N; i++) s += d.qr[j][i] * d.qr[j][i];
                    normsDirect[j] = normsUpdated[j] = sqrtf(s);
                } else normsUpdated[j] *= sqrtf(temp);
            }
        }
    }
    COD_UNROLL for (int k = 0; k < N; k++) d.perm[k] = k;
    COD_UNROLL for (int k = 0; k < N; k++) {
        COD_UNROLL for (int j = k + 1; j < N; j++) { const bool m = j == transp[k]; const int a = d.perm[k], b = d.perm[j]; d.perm[k] = m ? b : a; d.perm[j] = m ? a : b; }
    }
    const float premultiplied = fabsf(maxpivot) * (eps * (float)N);
    int rank = 0;
    COD_UNROLL for (int i = 0; i < N; i++) if (i < nonzero_pivots) rank += fabsf(d.qr[i][i]) > premultiplied;
    d.rank = rank;
}

// [R11 R12] -> [T11 0] for rank R < N (the second half of romis_cod_factor)
template <int N, int R> COD_HD void cod_fixed_z(CodFixed<N>& d) {
    COD_UNROLL for (int k = R - 1; k >= 0; --k) {
        if (k != R - 1) { COD_UNROLL for (int i = 0; i <= k; i++) { const float t = d.qr[k][i]; d.qr[k][i] = d.qr[R - 1][i]; d.qr[R - 1][i] = t; } }
        float tailSqNorm = 0.0f;
        COD_UNROLL for (int j = R; j < N; j++) tailSqNorm += d.qr[j][k] * d.qr[j][k];
        bool degenerate; float den, tau, beta;
        cod_householder_coeffs(d.qr[R - 1][k], tailSqNorm, degenerate, den, tau, beta);
        COD_UNROLL for (int j = R; j < N; j++) d.qr[j][k] = degenerate ? 0.0f : d.qr[j][k] / den;
        d.zc[k] = tau;
        d.qr[R - 1][k] = beta;
        if (k > 0 && tau != 0.0f) {
            COD_UNROLL for (int i = 0; i < k; i++) {
                float tmp = 0.0f;
                COD_UNROLL for (int j = R; j < N; j++) tmp += d.qr[j][i] * d.qr[j][k];
                tmp += d.qr[R - 1][i];
                d.qr[R - 1][i] -= tau * tmp;
                COD_UNROLL for (int j = R; j < N; j++) d.qr[j][i] -= (tau * tmp) * d.qr[j][k];
            }
        }
        if (k != R - 1) { COD_UNROLL for (int i = 0; i <= k; i++) { const float t = d.qr[k][i]; d.qr[k][i] = d.qr[R - 1][i]; d.qr[R - 1][i] = t; } }
    }
}

// romis_cod_solve for rank R
template <int N, int R> COD_HD void cod_fixed_solve(const CodFixed<N>& d, const float* b, float* x) {
    if (R == 0) { COD_UNROLL for (int i = 0; i < N; i++) x[i] = 0.0f; return; }
    float c[N];
    COD_UNROLL for (int i = 0; i < N; i++) c[i] = b[i];
    COD_UNROLL for (int k = 0; k < R; k++) {
        const float tau = d.hc[k];
        if (N - k > 1 && tau != 0.0f) {
            float tmp = 0.0f;
            COD_UNROLL for (int i = k + 1; i < N; i++) tmp += d.qr[k][i] * c[i];
            tmp += c[k];
            c[k] -= tau * tmp;
            COD_UNROLL for (int i = k + 1; i < N; i++) c[i] -= (tau * d.qr[k][i]) * tmp;
        }
    }
    COD_UNROLL for (int i = R - 1; i >= 0; --i) {
        if (c[i] != 0.0f) {
            c[i] /= d.qr[i][i];
            COD_UNROLL for (int s = 0; s < i; s++) c[s] -= c[i] * d.qr[i][s];
        }
    }
    if (R < N) {
        COD_UNROLL for (int i = R; i < N; i++) c[i] = 0.0f;
        COD_UNROLL for (int k = 0; k < R; k++) {
            if (k != R - 1) { const float t = c[k]; c[k] = c[R - 1]; c[R - 1] = t; }
            const float tau = d.zc[k];
            if (tau != 0.0f) {
                float tmp = 0.0f;
                COD_UNROLL for (int j = R; j < N; j++) tmp += d.qr[j][k] * c[j];
                tmp += c[R - 1];
                c[R - 1] -= tau * tmp;
                COD_UNROLL for (int j = R; j < N; j++) c[j] -= (tau * d.qr[j][k]) * tmp;
            }
            if (k != R - 1) { const float t = c[k]; c[k] = c[R - 1]; c[R - 1] = t; }
        }
    }
    COD_UNROLL for (int p = 0; p < N; p++) {        // perm is a permutation: every x[p] is taken exactly once
        float v = 0.0f;
        COD_UNROLL for (int i = 0; i < N; i++) v = d.perm[i] == p ? c[i] : v;
        x[p] = v;
    }
}

// One pixel's three solves: d.qr filled by the caller (symmetric technique matrix), load_b(ch, b) fetches a right-hand side,
// store_x(ch, x) takes the solution.  The rank switch is the only data-dependent branch.
template <int N, int R, class LoadB, class StoreX> COD_HD void cod_fixed_finish(CodFixed<N>& d, LoadB& load_b, StoreX& store_x) {
    if (R > 0 && R < N) cod_fixed_z<N, (R > 0 && R < N) ? R : 1>(d);
    for (int ch = 0; ch < 3; ch++) {
        float b[N], x[N];
        load_b(ch, b);
        cod_fixed_solve<N, R>(d, b, x);
        store_x(ch, x);
    }
}
template <int N, int R, class LoadB, class StoreX> struct CodDispatch {
    COD_HDM static void run(CodFixed<N>& d, LoadB& load_b, StoreX& store_x) {
        if (d.rank == R) cod_fixed_finish<N, R>(d, load_b, store_x);
        else CodDispatch<N, R - 1, LoadB, StoreX>::run(d, load_b, store_x);
    }
};
template <int N, class LoadB, class StoreX> struct CodDispatch<N, 0, LoadB, StoreX> {
    COD_HDM static void run(CodFixed<N>& d, LoadB& load_b, StoreX& store_x) { cod_fixed_finish<N, 0>(d, load_b, store_x); }
};
template <int N, class LoadB, class StoreX> COD_HD void cod_fixed_solve3(CodFixed<N>& d, LoadB load_b, StoreX store_x) {
    cod_fixed_qr<N>(d);
    CodDispatch<N, N, LoadB, StoreX>::run(d, load_b, store_x);
}

}  // namespace romis
