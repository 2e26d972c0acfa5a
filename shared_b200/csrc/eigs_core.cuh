// eigs_core.cuh -- the arithmetic of ArrayKernel.eigs, written once for the
// device and for a host emulation of one CTA.
//
// Replaces native/src/shared/LinearAlgebraOpsEigs.cpp:26-615 (Java twin: LinearAlgebraOps.java:462-), i.e. JAMA's
// nonsymmetric path: Householder reduction to Hessenberg form (orthes / ortran), Francis double-shift QR to the real
// Schur form (hqr2), back-substitution and back-transformation. The reference runs it as one scalar loop nest; here
// the same arithmetic is cut into phases of one cooperative thread array:
//   * whatever is a chain of dependent scalars (deflation search, shifts, the search for two small sub-diagonal
//     elements, the Householder vector of a column) is done by thread 0 and broadcast through EigsShared;
//   * every similarity update is done with ONE THREAD PER COLUMN (row modifications) or ONE THREAD PER ROW (column
//     modifications, accumulation into V), each thread walking its column/row in the reference's order.
// So every matrix element sees exactly the reference's sequence of IEEE operations (the file is compiled without FMA
// contraction): the result is bit-identical to LinearAlgebraOpsEigs.cpp, which the parity tests assert.
//   * back-substitution: the eigenvector of eigenvalue n only reads the Schur form's columns <= n and writes column n
//     (n-1 and n for a complex pair), so with the Schur form T kept read-only and the vectors written to a second
//     matrix X all columns are independent: one thread per eigenvalue (eigs_backsub_column);
//   * back-transformation vec = V X (upper triangular X): one thread per output element.
//
// The Cta parameter supplies tid / nthreads / sync(): threadIdx.x / blockDim.x / __syncthreads() on the device
// (array_eigs.cu), a pthread barrier in tests/emu/eigs_emu.cpp, where the same code runs under ThreadSanitizer so
// that a missing barrier shows up as a data race on the CPU.
#pragma once

#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define SSTC_HD __host__ __device__
#else
#define SSTC_HD
#endif

namespace sstc {

enum : int { EIGS_DONE = 0, EIGS_REAL_PAIR = 1, EIGS_QR_SWEEP = 2, EIGS_FAILED = 3 };

// hqr2 as the reference has it never gives up ("Could check iteration count here", LinearAlgebraOpsEigs.cpp:330): on
// an all-zero matrix of size >= 3 or on NaN input it spins forever. A kernel must not, so a root that has not
// converged after this many sweeps ends the call with an error (EISPACK's own limit is 30; JAMA's exceptional shifts
// fire at 10 and 30, so no converging input comes near).
constexpr int kEigsMaxSweepsPerRoot = 1000;

// Broadcast slots: written by thread 0 in its scalar phase, read by everybody after the next barrier.
struct EigsShared {
    double p, q, r, x;  // rotation / first Householder vector of a sweep, the shift x (hqr2's stale-x test)
    double hacc;        // orthes: the Householder scalar h; 0 when the column is already zero
    int mode, l, m, n;
};

SSTC_HD inline double eigs_eps() { return 2.220446049250313e-16; }  // pow(2.0, -52.0)

// LinearAlgebraOpsEigs.cpp:65-163. H: size x size with row stride ld (working copy of the source), V: size x size
// row-major (receives the accumulated transformations), ort: size doubles of scratch.
template <class Cta>
SSTC_HD inline void eigs_hessenberg(Cta &cta, double *H, int ld, double *V, double *ort, int size, EigsShared *sh) {
    const int tid = cta.tid, nt = cta.nthreads;
    const int high = size - 1;
    for (int i = tid; i < size; i += nt) {
        ort[i] = 0.0;
    }
    cta.sync();
    double scale = 0.0, g = 0.0;  // thread 0 only
    for (int m = 1; m <= high - 1; m++) {
        if (tid == 0) {
            scale = 0.0;
            for (int i = m; i <= high; i++) {
                scale = scale + fabs(H[i * ld + (m - 1)]);
            }
            double hacc = 0.0;
            if (scale != 0.0) {
                for (int i = high; i >= m; i--) {
                    ort[i] = H[i * ld + (m - 1)] / scale;
                    hacc += ort[i] * ort[i];
                }
                g = sqrt(hacc);
                if (ort[m] > 0) {
                    g = -g;
                }
                hacc = hacc - ort[m] * g;
                ort[m] = ort[m] - g;
            }
            sh->hacc = hacc;
            sh->mode = scale != 0.0;
        }
        cta.sync();
        if (sh->mode) {  // uniform
            const double hacc = sh->hacc;
            // H = (I - u u'/h) H : one thread per column j
            for (int j = m + tid; j < size; j += nt) {
                double f = 0.0;
                for (int i = high; i >= m; i--) {
                    f += ort[i] * H[i * ld + j];
                }
                f = f / hacc;
                for (int i = m; i <= high; i++) {
                    H[i * ld + j] -= f * ort[i];
                }
            }
            cta.sync();
            // H = H (I - u u'/h) : one thread per row i
            for (int i = tid; i <= high; i += nt) {
                double f = 0.0;
                for (int j = high; j >= m; j--) {
                    f += ort[j] * H[i * ld + j];
                }
                f = f / hacc;
                for (int j = m; j <= high; j++) {
                    H[i * ld + j] -= f * ort[j];
                }
            }
            cta.sync();
            if (tid == 0) {
                ort[m] = scale * ort[m];
                H[m * ld + (m - 1)] = scale * g;
            }
        } else {
            cta.sync();  // nobody still reads *sh when thread 0 starts the next column
        }
        // the next scalar phase is thread 0's again; everybody else waits at its barrier
    }
    cta.sync();
    // ortran: V = I, then the reflectors from the last to the first
    for (int e = tid; e < size * size; e += nt) {
        V[e] = (e / size == e % size) ? 1.0 : 0.0;
    }
    cta.sync();
    for (int m = high - 1; m >= 1; m--) {
        const double sub = H[m * ld + (m - 1)];  // H is read-only from here on: uniform
        if (sub != 0.0) {
            for (int i = m + 1 + tid; i <= high; i += nt) {
                ort[i] = H[i * ld + (m - 1)];
            }
            cta.sync();
            for (int j = m + tid; j <= high; j += nt) {
                double gg = 0.0;
                for (int i = m; i <= high; i++) {
                    gg += ort[i] * V[i * size + j];
                }
                gg = (gg / ort[m]) / sub;  // double division avoids possible underflow
                for (int i = m; i <= high; i++) {
                    V[i * size + j] += gg * ort[i];
                }
            }
            cta.sync();
        }
    }
}

// LinearAlgebraOpsEigs.cpp:165-453 (hqr2 up to, not including, the back-substitution). On return H holds the real
// Schur form, V the accumulated transformations, val the interleaved (re, im) eigenvalues; *norm_out the norm
// hqr2 computes first (thread 0 writes it). sh->mode is EIGS_DONE or EIGS_FAILED afterwards.
template <class Cta>
SSTC_HD inline void eigs_schur(Cta &cta, double *H, int ld, double *V, double *val, int nn, EigsShared *sh,
        double *norm_out) {
    const int tid = cta.tid, nt = cta.nthreads;
    const double eps = eigs_eps();
    // thread 0's private state
    int n = nn - 1, iter = 0;
    double exshift = 0.0, norm = 0.0;
    if (tid == 0) {
        for (int i = 0; i < nn; i++) {
            for (int j = (i - 1 > 0 ? i - 1 : 0); j < nn; j++) {
                norm = norm + fabs(H[i * ld + j]);
            }
        }
        *norm_out = norm;
    }
    for (;;) {
        if (tid == 0) {
            int mode = EIGS_DONE;
            while (n >= 0) {
                double s, w, x, y, z, p, q, r;
                int l = n;
                while (l > 0) {
                    s = fabs(H[(l - 1) * ld + (l - 1)]) + fabs(H[l * ld + l]);
                    if (s == 0.0) {
                        s = norm;
                    }
                    if (fabs(H[l * ld + (l - 1)]) < eps * s) {
                        break;
                    }
                    l--;
                }
                if (l == n) {  // one root found
                    H[n * ld + n] = H[n * ld + n] + exshift;
                    val[2 * n] = H[n * ld + n];
                    val[2 * n + 1] = 0.0;
                    n--;
                    iter = 0;
                    continue;
                }
                if (l == n - 1) {  // two roots found
                    w = H[n * ld + (n - 1)] * H[(n - 1) * ld + n];
                    p = (H[(n - 1) * ld + (n - 1)] - H[n * ld + n]) / 2.0;
                    q = p * p + w;
                    z = sqrt(fabs(q));
                    H[n * ld + n] = H[n * ld + n] + exshift;
                    H[(n - 1) * ld + (n - 1)] = H[(n - 1) * ld + (n - 1)] + exshift;
                    x = H[n * ld + n];
                    if (q >= 0) {  // real pair: a plane rotation of rows / columns n-1, n follows
                        if (p >= 0) {
                            z = p + z;
                        } else {
                            z = p - z;
                        }
                        val[2 * (n - 1)] = x + z;
                        val[2 * n] = val[2 * (n - 1)];
                        if (z != 0.0) {
                            val[2 * n] = x - w / z;
                        }
                        val[2 * (n - 1) + 1] = 0.0;
                        val[2 * n + 1] = 0.0;
                        x = H[n * ld + (n - 1)];
                        s = fabs(x) + fabs(z);
                        p = x / s;
                        q = z / s;
                        r = sqrt(p * p + q * q);
                        sh->p = p / r;
                        sh->q = q / r;
                        sh->n = n;
                        mode = EIGS_REAL_PAIR;
                        n = n - 2;
                        iter = 0;
                        break;
                    }
                    val[2 * (n - 1)] = x + p;  // complex pair
                    val[2 * n] = x + p;
                    val[2 * (n - 1) + 1] = z;
                    val[2 * n + 1] = -z;
                    n = n - 2;
                    iter = 0;
                    continue;
                }
                // no convergence yet: form the shift
                x = H[n * ld + n];
                y = H[(n - 1) * ld + (n - 1)];  // l < n always holds here
                w = H[n * ld + (n - 1)] * H[(n - 1) * ld + n];
                if (iter == 10) {  // Wilkinson's original ad hoc shift
                    exshift += x;
                    for (int i = 0; i <= n; i++) {
                        H[i * ld + i] -= x;
                    }
                    s = fabs(H[n * ld + (n - 1)]) + fabs(H[(n - 1) * ld + (n - 2)]);
                    x = y = 0.75 * s;
                    w = -0.4375 * s * s;
                }
                if (iter == 30) {  // MATLAB's new ad hoc shift
                    s = (y - x) / 2.0;
                    s = s * s + w;
                    if (s > 0) {
                        s = sqrt(s);
                        if (y < x) {
                            s = -s;
                        }
                        s = x - w / ((y - x) / 2.0 + s);
                        for (int i = 0; i <= n; i++) {
                            H[i * ld + i] -= s;
                        }
                        exshift += s;
                        x = y = w = 0.964;
                    }
                }
                iter = iter + 1;
                if (iter > kEigsMaxSweepsPerRoot) {
                    mode = EIGS_FAILED;
                    break;
                }
                // look for two consecutive small sub-diagonal elements
                int m = n - 2;
                p = q = r = 0.0;
                while (m >= l) {
                    z = H[m * ld + m];
                    r = x - z;
                    s = y - z;
                    p = (r * s - w) / H[(m + 1) * ld + m] + H[m * ld + (m + 1)];
                    q = H[(m + 1) * ld + (m + 1)] - z - r - s;
                    r = H[(m + 2) * ld + (m + 1)];
                    s = fabs(p) + fabs(q) + fabs(r);
                    p = p / s;
                    q = q / s;
                    r = r / s;
                    if (m == l) {
                        break;
                    }
                    if (fabs(H[m * ld + (m - 1)]) * (fabs(q) + fabs(r)) <
                            eps * (fabs(p) * (fabs(H[(m - 1) * ld + (m - 1)]) + fabs(z) +
                                                     fabs(H[(m + 1) * ld + (m + 1)])))) {
                        break;
                    }
                    m--;
                }
                for (int i = m + 2; i <= n; i++) {
                    H[i * ld + (i - 2)] = 0.0;
                    if (i > m + 2) {
                        H[i * ld + (i - 3)] = 0.0;
                    }
                }
                sh->p = p;
                sh->q = q;
                sh->r = r;
                sh->x = x;
                sh->l = l;
                sh->m = m;
                sh->n = n;
                mode = EIGS_QR_SWEEP;
                break;
            }
            sh->mode = mode;
        }
        cta.sync();
        const int mode = sh->mode;
        if (mode == EIGS_DONE || mode == EIGS_FAILED) {
            break;  // sh->mode keeps the outcome for the caller
        }
        if (mode == EIGS_REAL_PAIR) {
            const int n2 = sh->n;
            const double p = sh->p, q = sh->q;
            for (int j = n2 - 1 + tid; j < nn; j += nt) {  // row modification
                const double z = H[(n2 - 1) * ld + j];
                H[(n2 - 1) * ld + j] = q * z + p * H[n2 * ld + j];
                H[n2 * ld + j] = q * H[n2 * ld + j] - p * z;
            }
            cta.sync();
            for (int i = tid; i < nn; i += nt) {
                if (i <= n2) {  // column modification
                    const double z = H[i * ld + (n2 - 1)];
                    H[i * ld + (n2 - 1)] = q * z + p * H[i * ld + n2];
                    H[i * ld + n2] = q * H[i * ld + n2] - p * z;
                }
                const double z = V[i * nn + (n2 - 1)];  // accumulate transformations
                V[i * nn + (n2 - 1)] = q * z + p * V[i * nn + n2];
                V[i * nn + n2] = q * V[i * nn + n2] - p * z;
            }
            cta.sync();
            continue;
        }
        // double QR step involving rows l:n and columns m:n; every thread tracks the scalars redundantly
        {
            const int l = sh->l, m = sh->m, n2 = sh->n;
            double p = sh->p, q = sh->q, r = sh->r, x = sh->x, y, z, s;
            for (int k = m; k <= n2 - 1; k++) {
                const bool not_last = (k != n2 - 1);
                if (k != m) {
                    p = H[k * ld + (k - 1)];
                    q = H[(k + 1) * ld + (k - 1)];
                    r = (not_last ? H[(k + 2) * ld + (k - 1)] : 0.0);
                    x = fabs(p) + fabs(q) + fabs(r);
                    if (x != 0.0) {
                        p = p / x;
                        q = q / x;
                        r = r / x;
                    }
                }
                if (x == 0.0) {
                    break;
                }
                s = sqrt(p * p + q * q + r * r);
                if (p < 0) {
                    s = -s;
                }
                if (s != 0) {
                    const double sub = -s * x;  // H[k][k-1] for k != m, stored by thread 0 after the barrier
                    p = p + s;
                    x = p / s;
                    y = q / s;
                    z = r / s;
                    q = q / p;
                    r = r / p;
                    for (int j = k + tid; j < nn; j += nt) {  // row modification
                        double t = H[k * ld + j] + q * H[(k + 1) * ld + j];
                        if (not_last) {
                            t = t + r * H[(k + 2) * ld + j];
                            H[(k + 2) * ld + j] = H[(k + 2) * ld + j] - t * z;
                        }
                        H[k * ld + j] = H[k * ld + j] - t * x;
                        H[(k + 1) * ld + j] = H[(k + 1) * ld + j] - t * y;
                    }
                    cta.sync();
                    if (tid == 0) {
                        if (k != m) {
                            H[k * ld + (k - 1)] = sub;
                        } else if (l != m) {
                            H[k * ld + (k - 1)] = -H[k * ld + (k - 1)];
                        }
                    }
                    const int ilast = n2 < k + 3 ? n2 : k + 3;
                    for (int i = tid; i < nn; i += nt) {
                        if (i <= ilast) {  // column modification
                            double t = x * H[i * ld + k] + y * H[i * ld + (k + 1)];
                            if (not_last) {
                                t = t + z * H[i * ld + (k + 2)];
                                H[i * ld + (k + 2)] = H[i * ld + (k + 2)] - t * r;
                            }
                            H[i * ld + k] = H[i * ld + k] - t;
                            H[i * ld + (k + 1)] = H[i * ld + (k + 1)] - t * q;
                        }
                        double t = x * V[i * nn + k] + y * V[i * nn + (k + 1)];  // accumulate transformations
                        if (not_last) {
                            t = t + z * V[i * nn + (k + 2)];
                            V[i * nn + (k + 2)] = V[i * nn + (k + 2)] - t * r;
                        }
                        V[i * nn + k] = V[i * nn + k] - t;
                        V[i * nn + (k + 1)] = V[i * nn + (k + 1)] - t * q;
                    }
                    cta.sync();
                }
            }
            cta.sync();  // nobody still reads *sh when thread 0 starts its next scalar phase
        }
    }
}

SSTC_HD inline void eigs_cdiv(double are, double aim, double bre, double bim, double &re, double &im) {
    // Common.hpp:344-348 (jcomplex operator/)
    re = (are * bre + aim * bim) / (bre * bre + bim * bim);
    im = (aim * bre - are * bim) / (bre * bre + bim * bim);
}

// LinearAlgebraOpsEigs.cpp:462-592 for ONE eigenvalue index n: T is the Schur form (row-major nn x nn, read-only),
// X receives column n (and n-1 for the second member of a complex pair). The first member of a pair does nothing.
SSTC_HD inline void eigs_backsub_column(const double *T, double *X, const double *val, int nn, int n, double norm) {
    const double eps = eigs_eps();
    const double p = val[2 * n];
    double q = val[2 * n + 1];
    double r = 0.0, s = 0.0, z = 0.0, t, w, x, y;
    if (q == 0) {  // real vector
        int l = n;
        X[n * nn + n] = 1.0;
        for (int i = n - 1; i >= 0; i--) {
            w = T[i * nn + i] - p;
            r = 0.0;
            for (int j = l; j <= n; j++) {
                r = r + T[i * nn + j] * X[j * nn + n];
            }
            if (val[2 * i + 1] < 0.0) {
                z = w;
                s = r;
            } else {
                l = i;
                if (val[2 * i + 1] == 0.0) {
                    if (w != 0.0) {
                        X[i * nn + n] = -r / w;
                    } else {
                        X[i * nn + n] = -r / (eps * norm);
                    }
                } else {  // solve real equations
                    x = T[i * nn + (i + 1)];
                    y = T[(i + 1) * nn + i];
                    q = (val[2 * i] - p) * (val[2 * i] - p) + val[2 * i + 1] * val[2 * i + 1];
                    t = (x * s - z * r) / q;
                    X[i * nn + n] = t;
                    if (fabs(x) > fabs(z)) {
                        X[(i + 1) * nn + n] = (-r - w * t) / x;
                    } else {
                        X[(i + 1) * nn + n] = (-s - y * t) / z;
                    }
                }
                t = fabs(X[i * nn + n]);  // overflow control
                if ((eps * t) * t > 1) {
                    for (int j = i; j <= n; j++) {
                        X[j * nn + n] = X[j * nn + n] / t;
                    }
                }
            }
        }
    } else if (q < 0) {  // complex vector: columns n-1 (real part) and n (imaginary part)
        int l = n - 1;
        double cre, cim;
        if (fabs(T[n * nn + (n - 1)]) > fabs(T[(n - 1) * nn + n])) {
            X[(n - 1) * nn + (n - 1)] = q / T[n * nn + (n - 1)];
            X[(n - 1) * nn + n] = -(T[n * nn + n] - p) / T[n * nn + (n - 1)];
        } else {
            eigs_cdiv(0.0, -T[(n - 1) * nn + n], T[(n - 1) * nn + (n - 1)] - p, q, cre, cim);
            X[(n - 1) * nn + (n - 1)] = cre;
            X[(n - 1) * nn + n] = cim;
        }
        X[n * nn + (n - 1)] = 0.0;
        X[n * nn + n] = 1.0;
        for (int i = n - 2; i >= 0; i--) {
            double ra = 0.0, sa = 0.0, vr, vi;
            for (int j = l; j <= n; j++) {
                ra = ra + T[i * nn + j] * X[j * nn + (n - 1)];
                sa = sa + T[i * nn + j] * X[j * nn + n];
            }
            w = T[i * nn + i] - p;
            if (val[2 * i + 1] < 0.0) {
                z = w;
                r = ra;
                s = sa;
            } else {
                l = i;
                if (val[2 * i + 1] == 0) {
                    eigs_cdiv(-ra, -sa, w, q, cre, cim);
                    X[i * nn + (n - 1)] = cre;
                    X[i * nn + n] = cim;
                } else {  // solve complex equations
                    x = T[i * nn + (i + 1)];
                    y = T[(i + 1) * nn + i];
                    vr = (val[2 * i] - p) * (val[2 * i] - p) + val[2 * i + 1] * val[2 * i + 1] - q * q;
                    vi = (val[2 * i] - p) * 2.0 * q;
                    if (vr == 0.0 && vi == 0.0) {
                        vr = eps * norm * (fabs(w) + fabs(q) + fabs(x) + fabs(y) + fabs(z));
                    }
                    eigs_cdiv(x * r - z * ra + q * sa, x * s - z * sa - q * ra, vr, vi, cre, cim);
                    X[i * nn + (n - 1)] = cre;
                    X[i * nn + n] = cim;
                    if (fabs(x) > (fabs(z) + fabs(q))) {
                        X[(i + 1) * nn + (n - 1)] = (-ra - w * X[i * nn + (n - 1)] + q * X[i * nn + n]) / x;
                        X[(i + 1) * nn + n] = (-sa - w * X[i * nn + n] - q * X[i * nn + (n - 1)]) / x;
                    } else {
                        eigs_cdiv(-r - y * X[i * nn + (n - 1)], -s - y * X[i * nn + n], z, q, cre, cim);
                        X[(i + 1) * nn + (n - 1)] = cre;
                        X[(i + 1) * nn + n] = cim;
                    }
                }
                const double ta = fabs(X[i * nn + (n - 1)]), tb = fabs(X[i * nn + n]);  // overflow control
                t = ta < tb ? tb : ta;  // std::max(a, b)
                if ((eps * t) * t > 1) {
                    for (int j = i; j <= n; j++) {
                        X[j * nn + (n - 1)] = X[j * nn + (n - 1)] / t;
                        X[j * nn + n] = X[j * nn + n] / t;
                    }
                }
            }
        }
    }
}

// LinearAlgebraOpsEigs.cpp:604-613 for ONE output element: vec[i][j] = sum_{k <= j} V[i][k] X[k][j], k ascending.
SSTC_HD inline double eigs_backtransform_element(const double *V, const double *X, int nn, int i, int j) {
    double z = 0.0;
    for (int k = 0; k <= j; k++) {
        z = z + V[i * nn + k] * X[k * nn + j];
    }
    return z;
}

}  // namespace sstc
