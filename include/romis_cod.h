/*
 * romis_cod.h -- `A.completeOrthogonalDecomposition().solve(b)` for one small dense float system, as R-OMIS needs it
 * per pixel (solveSystem, reference src/rendering/render_utils.h:52; Eigen is vendored under reference src/Eigen).
 *
 * Restated from the published algorithm of Eigen 3.4's CompleteOrthogonalDecomposition / ColPivHouseholderQR
 * (src/Eigen/src/QR/ColPivHouseholderQR.h:481-577 computeInPlace, :255-264 rank;
 *  src/Eigen/src/QR/CompleteOrthogonalDecomposition.h computeInPlace / _solve_impl / applyZAdjointOnTheLeftInPlace;
 *  src/Eigen/src/Householder/Householder.h makeHouseholder / applyHouseholderOnTheLeft / applyHouseholderOnTheRight):
 *   1. Householder QR with column pivoting: pivot = column of largest running norm (first maximum), LAPACK-style norm
 *      downdating with recomputation when the update has lost accuracy (lawn176), numerically-zero pivots tracked;
 *   2. rank = number of diagonal entries with |R_ii| > |max pivot| * epsilon * n;
 *   3. for rank < n, Householder reflections from the right reduce [R11 R12] to [T11 0] (the "Z" factor);
 *   4. solve: c = Q^T b, back substitution with T11, zero padding, Z^T, column permutation
 *      -- the minimum-norm least-squares solution.
 * Every reduction (column norms, Householder dot products) is summed front to back in fp32.  Eigen sums them in SSE
 * packets whose grouping depends on the alignment of each sub-vector, which is not reproducible outside Eigen, so the
 * result agrees with the reference's to rounding, not to the bit; what IS bit-exact is this routine on the CPU (oracle,
 * -ffp-contract=off) against this routine on the GPU (-fmad=false): one definition, like romis_detmath.h.
 */
#ifndef ROMIS_COD_H
#define ROMIS_COD_H

#include <float.h>
#include <math.h>

#if defined(__CUDACC__)
#define ROMIS_COD_HD __host__ __device__ __forceinline__
#else
#define ROMIS_COD_HD static inline
#endif

#define ROMIS_COD_MAX 11        /* numNeighboursToSample + 1 <= 11 (ui.cpp:307: k <= 10) */

typedef struct romis_cod {
    int n, rank;
    int perm[ROMIS_COD_MAX];                    /* colsPermutation().indices() */
    float qr[ROMIS_COD_MAX * ROMIS_COD_MAX];    /* column-major with leading dimension n (the first n*n floats are used: a thread of the
                                                   GPU solve touches 144 bytes for k = 5 instead of 484): R / T above the diagonal,
                                                   Householder essentials below and right */
    float hc[ROMIS_COD_MAX], zc[ROMIS_COD_MAX]; /* hCoeffs, zCoeffs */
} romis_cod;

#define ROMIS_QR(d, i, j) ((d)->qr[(j) * ld + (i)])     /* `ld` = the system's n, a local of every user */

/* makeHouseholder on the vector {*c0, tail[0 .. len*stride)} (Householder.h:67-96): on return *c0 is untouched, the tail
 * holds the essential part, *tau and *beta are set. */
ROMIS_COD_HD void romis_make_householder(const float* c0p, float* tail, int len, int stride, float* tau, float* beta) {
    float tailSqNorm = 0.0f;
    for (int i = 0; i < len; i++) tailSqNorm += tail[i * stride] * tail[i * stride];
    const float c0 = *c0p;
    if (tailSqNorm <= FLT_MIN) {
        *tau = 0.0f; *beta = c0;
        for (int i = 0; i < len; i++) tail[i * stride] = 0.0f;
    } else {
        float b = sqrtf(c0 * c0 + tailSqNorm);
        if (c0 >= 0.0f) b = -b;
        const float den = c0 - b;
        for (int i = 0; i < len; i++) tail[i * stride] = tail[i * stride] / den;
        *tau = (b - c0) / b;
        *beta = b;
    }
}

/* The decomposition of the n x n matrix the caller has put into d->qr (column-major, leading dimension n), in place. */
ROMIS_COD_HD void romis_cod_factor(romis_cod* d, int n) {
    const float eps = FLT_EPSILON;
    const int ld = n;
    float normsUpdated[ROMIS_COD_MAX], normsDirect[ROMIS_COD_MAX];
    int transp[ROMIS_COD_MAX];
    d->n = n;
    float maxNorm = 0.0f;
    for (int k = 0; k < n; k++) {
        float s = 0.0f;
        for (int i = 0; i < n; i++) s += ROMIS_QR(d, i, k) * ROMIS_QR(d, i, k);
        normsDirect[k] = normsUpdated[k] = sqrtf(s);
        if (k == 0 || normsUpdated[k] > maxNorm) maxNorm = normsUpdated[k];
    }
    const float me = maxNorm * eps;
    const float threshold_helper = (me * me) / (float)n;
    const float norm_downdate_threshold = sqrtf(eps);
    int nonzero_pivots = n;
    float maxpivot = 0.0f;
    for (int k = 0; k < n; k++) {
        int big = k; float bigNorm = normsUpdated[k];
        for (int j = k + 1; j < n; j++) if (normsUpdated[j] > bigNorm) { bigNorm = normsUpdated[j]; big = j; }
        if (nonzero_pivots == n && bigNorm * bigNorm < threshold_helper * (float)(n - k)) nonzero_pivots = k;
        transp[k] = big;
        if (k != big) {
            for (int i = 0; i < n; i++) { float t = ROMIS_QR(d, i, k); ROMIS_QR(d, i, k) = ROMIS_QR(d, i, big); ROMIS_QR(d, i, big) = t; }
            float t = normsUpdated[k]; normsUpdated[k] = normsUpdated[big]; normsUpdated[big] = t;
            t = normsDirect[k]; normsDirect[k] = normsDirect[big]; normsDirect[big] = t;
        }
        float tau, beta;
        romis_make_householder(&ROMIS_QR(d, k, k), &ROMIS_QR(d, k + 1, k), n - k - 1, 1, &tau, &beta);
        d->hc[k] = tau;
        ROMIS_QR(d, k, k) = beta;
        if (fabsf(beta) > maxpivot) maxpivot = fabsf(beta);
        /* bottomRightCorner(n-k, n-k-1).applyHouseholderOnTheLeft(essential, tau)  (Householder.h:116-134) */
        if (n - k > 1 && tau != 0.0f) {
            for (int j = k + 1; j < n; j++) {
                float tmp = 0.0f;
                for (int i = k + 1; i < n; i++) tmp += ROMIS_QR(d, i, k) * ROMIS_QR(d, i, j);
                tmp += ROMIS_QR(d, k, j);
                ROMIS_QR(d, k, j) -= tau * tmp;
                for (int i = k + 1; i < n; i++) ROMIS_QR(d, i, j) -= (tau * ROMIS_QR(d, i, k)) * tmp;
            }
        }
        for (int j = k + 1; j < n; j++) {                       /* norm downdate (ColPivHouseholderQR.h:548-567) */
            if (normsUpdated[j] != 0.0f) {
                float temp = fabsf(ROMIS_QR(d, k, j)) / normsUpdated[j];
                temp = (1.0f + temp) * (1.0f - temp);
                temp = temp < 0.0f ? 0.0f : temp;
                const float ratio = normsUpdated[j] / normsDirect[j];
                const float temp2 = temp * (ratio * ratio);
                if (temp2 <= norm_downdate_threshold) {
                    float s = 0.0f;
                    for (int i = k + 1; i < n; i++) s += ROMIS_QR(d, i, j) * ROMIS_QR(d, i, j);
                    normsDirect[j] = normsUpdated[j] = sqrtf(s);
                } else normsUpdated[j] *= sqrtf(temp);
            }
        }
    }
    for (int k = 0; k < n; k++) d->perm[k] = k;
    for (int k = 0; k < n; k++) { int t = d->perm[k]; d->perm[k] = d->perm[transp[k]]; d->perm[transp[k]] = t; }
    /* rank (ColPivHouseholderQR.h:255-264), default threshold = epsilon * diagonalSize */
    const float premultiplied = fabsf(maxpivot) * (eps * (float)n);
    int rank = 0;
    for (int i = 0; i < nonzero_pivots; i++) rank += fabsf(ROMIS_QR(d, i, i)) > premultiplied;
    d->rank = rank;
    /* [R11 R12] -> [T11 0] by reflections from the right (CompleteOrthogonalDecomposition::computeInPlace) */
    if (rank < n) {
        for (int k = rank - 1; k >= 0; --k) {
            if (k != rank - 1)
                for (int i = 0; i <= k; i++) { float t = ROMIS_QR(d, i, k); ROMIS_QR(d, i, k) = ROMIS_QR(d, i, rank - 1); ROMIS_QR(d, i, rank - 1) = t; }
            float tau, beta;
            romis_make_householder(&ROMIS_QR(d, k, rank - 1), &ROMIS_QR(d, k, rank), n - rank, ld, &tau, &beta);
            d->zc[k] = tau;
            ROMIS_QR(d, k, rank - 1) = beta;
            /* topRightCorner(k, n-rank+1).applyHouseholderOnTheRight(row(k).tail(n-rank)^T, tau)  (Householder.h:153-171) */
            if (k > 0 && tau != 0.0f) {
                for (int i = 0; i < k; i++) {
                    float tmp = 0.0f;
                    for (int j = rank; j < n; j++) tmp += ROMIS_QR(d, i, j) * ROMIS_QR(d, k, j);
                    tmp += ROMIS_QR(d, i, rank - 1);
                    ROMIS_QR(d, i, rank - 1) -= tau * tmp;
                    for (int j = rank; j < n; j++) ROMIS_QR(d, i, j) -= (tau * tmp) * ROMIS_QR(d, k, j);
                }
            }
            if (k != rank - 1)
                for (int i = 0; i <= k; i++) { float t = ROMIS_QR(d, i, k); ROMIS_QR(d, i, k) = ROMIS_QR(d, i, rank - 1); ROMIS_QR(d, i, rank - 1) = t; }
        }
    }
}

/* A is row-major n x n */
ROMIS_COD_HD void romis_cod_compute(romis_cod* d, const float* A, int n) {
    const int ld = n;
    for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) ROMIS_QR(d, i, j) = A[i * n + j];
    romis_cod_factor(d, n);
}

/* x = minimum-norm least-squares solution of A x = b (CompleteOrthogonalDecomposition::_solve_impl) */
ROMIS_COD_HD void romis_cod_solve(const romis_cod* d, const float* b, float* x) {
    const int n = d->n > ROMIS_COD_MAX ? ROMIS_COD_MAX : d->n;
    const int ld = n;
    const int rank = d->rank < 0 ? 0 : (d->rank > n ? n : d->rank);
    float c[ROMIS_COD_MAX] = {0.0f};
    if (rank == 0) { for (int i = 0; i < n; i++) x[i] = 0.0f; return; }
    for (int i = 0; i < n; i++) c[i] = b[i];
    for (int k = 0; k < rank; k++) {                            /* c = Q^T b: H_0 first (HouseholderSequence, adjoint) */
        const float tau = d->hc[k];
        if (n - k > 1 && tau != 0.0f) {
            float tmp = 0.0f;
            for (int i = k + 1; i < n; i++) tmp += ROMIS_QR(d, i, k) * c[i];
            tmp += c[k];
            c[k] -= tau * tmp;
            for (int i = k + 1; i < n; i++) c[i] -= (tau * ROMIS_QR(d, i, k)) * tmp;
        }
    }
    for (int i = rank - 1; i >= 0; --i) {                       /* T11 y = c, column-oriented back substitution */
        if (c[i] != 0.0f) {
            c[i] /= ROMIS_QR(d, i, i);
            for (int s = 0; s < i; s++) c[s] -= c[i] * ROMIS_QR(d, s, i);
        }
    }
    if (rank < n) {
        for (int i = rank; i < n; i++) c[i] = 0.0f;
        for (int k = 0; k < rank; k++) {                        /* applyZAdjointOnTheLeftInPlace */
            if (k != rank - 1) { float t = c[k]; c[k] = c[rank - 1]; c[rank - 1] = t; }
            const float tau = d->zc[k];
            if (tau != 0.0f) {
                float tmp = 0.0f;
                for (int j = rank; j < n; j++) tmp += ROMIS_QR(d, k, j) * c[j];
                tmp += c[rank - 1];
                c[rank - 1] -= tau * tmp;
                for (int j = rank; j < n; j++) c[j] -= (tau * ROMIS_QR(d, k, j)) * tmp;
            }
            if (k != rank - 1) { float t = c[k]; c[k] = c[rank - 1]; c[rank - 1] = t; }
        }
    }
    for (int i = 0; i < n; i++) x[d->perm[i]] = c[i];
}

#endif /* ROMIS_COD_H */
