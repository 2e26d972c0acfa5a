// eigs_core.cuh -- the arithmetic of ArrayKernel.eigs, written once for the
// device and for a host emulation of one CTA.
//
// Replaces native/src/shared/LinearAlgebraOpsEigs.cpp:26-615 (Java twin: LinearAlgebraOps.java:462-), i.e. JAMA's
// nonsymmetric path: Householder reduction to Hessenberg form (orthes / ortran), Francis double-shift QR to the real
// Schur form (hqr2), back-substitution and back-transformation. The reference runs it as one scalar loop nest; here
// the same arithmetic is cut into phases of one cooperative thread array:
//   * chains of dependent scalars (the Householder vector of a column, root extraction, shifts) are done by thread 0
//     and broadcast through EigsShared;
//   * the two searches that open a QR sweep (a small sub-diagonal element; two consecutive small ones) are sets of
//     independent tests: every thread tests its own indices and the largest hit wins (atomic max);
//   * every similarity update of H is done with ONE THREAD PER COLUMN (row modifications) or ONE THREAD PER ROW
//     (column modifications), each thread walking its column/row in the reference's order;
//   * the QR iteration never reads V, so its steps are only RECORDED while it runs (EigsRot) and applied to V
//     afterwards with one thread per row of V walking the log front to back -- no barrier per step, and V can take
//     H's place in shared memory once H is done.
// So every matrix element sees exactly the reference's sequence of IEEE operations (the file is compiled without FMA
// contraction): the result is bit-identical to LinearAlgebraOpsEigs.cpp, which the parity tests assert.
//   * back-substitution: the eigenvector of eigenvalue n only reads the Schur form's columns <= n and writes column n
//     (n-1 and n for a complex pair), so with the Schur form T kept read-only and the vectors written to a second
//     matrix X all columns are independent: one thread per eigenvalue (eigs_backsub_column);
//   * back-transformation vec = V X (upper triangular X): one thread per output element.
//
// The Cta parameter supplies tid / nthreads / sync() / atomic_max(): threadIdx.x / blockDim.x / __syncthreads() /
// atomicMax() on the device
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

enum : int { EIGS_DONE = 0, EIGS_REAL_PAIR = 1, EIGS_QR_SWEEP = 2, EIGS_FAILED = 3, EIGS_NEXT = 4 };

// hqr2 as the reference has it never gives up ("Could check iteration count here", LinearAlgebraOpsEigs.cpp:330): on
// an all-zero matrix of size >= 3 or on NaN input it spins forever. A kernel must not, so a root that has not
// converged after this many sweeps ends the call with an error (EISPACK's own limit is 30; JAMA's exceptional shifts
// fire at 10 and 30, so no converging input comes near).
constexpr int kEigsMaxSweepsPerRoot = 1000;

// Broadcast slots: written by thread 0 in its scalar phase, read by everybody after the next barrier.
struct EigsShared {
    double p, q;        // plane rotation of a real pair
    double x, y, w;     // the shift of a sweep
    double norm;
    double hacc;        // orthes: the Householder scalar h; 0 when the column is already zero
    int mode, l, m, n;  // l, m: targets of the atomic-max searches; n: index of the root being isolated
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

// One similarity step of hqr2 as it acts on V: recorded by thread 0 while the QR iteration runs on H alone, applied to
// V afterwards (eigs_apply_rots). 48 bytes.
struct EigsRot {
    double x, y, z, q, r;  // EIGS_ROT_PLANE: x = p, y = q of the plane rotation
    int32_t k, kind;
};
enum : int { EIGS_ROT_QR3 = 0, EIGS_ROT_QR2 = 1, EIGS_ROT_PLANE = 2 };

// V <- V * (the recorded steps, in order): one thread per ROW of V, every row walking the log front to back, so each
// element sees the reference's operation order (LinearAlgebraOpsEigs.cpp:284-290, 437-447). The log is staged through
// `chunk` (shared memory on the device) chunk_cap entries at a time. Ends with a barrier.
template <class Cta>
SSTC_HD inline void eigs_apply_rots(Cta &cta, double *V, int ldv, int nn, const EigsRot *log, int count, EigsRot *chunk,
        int chunk_cap) {
    const int tid = cta.tid, nt = cta.nthreads;
    for (int base = 0; base < count; base += chunk_cap) {
        const int m = count - base < chunk_cap ? count - base : chunk_cap;
        const uint64_t *src = reinterpret_cast<const uint64_t *>(log + base);
        uint64_t *dst = reinterpret_cast<uint64_t *>(chunk);
        for (int w = tid; w < m * 6; w += nt) {
            dst[w] = src[w];
        }
        cta.sync();
        for (int i = tid; i < nn; i += nt) {
            double *row = V + (int64_t) i * ldv;
            for (int t = 0; t < m; t++) {
                const EigsRot &e = chunk[t];
                const int k = e.k;
                if (e.kind == EIGS_ROT_PLANE) {
                    const double z = row[k];
                    row[k] = e.y * z + e.x * row[k + 1];
                    row[k + 1] = e.y * row[k + 1] - e.x * z;
                } else {
                    double p = e.x * row[k] + e.y * row[k + 1];
                    if (e.kind == EIGS_ROT_QR3) {
                        p = p + e.z * row[k + 2];
                        row[k + 2] = row[k + 2] - p * e.r;
                    }
                    row[k] = row[k] - p;
                    row[k + 1] = row[k + 1] - p * e.q;
                }
            }
        }
        cta.sync();
    }
}

// LinearAlgebraOpsEigs.cpp:165-453 (hqr2 up to, not including, the back-substitution), on H alone. On return H holds
// the real Schur form, val the interleaved (re, im) eigenvalues, *norm_out the norm hqr2 computes first, and log[0 ..
// return value) the steps still to be applied to V (eigs_apply_rots); when the log fills up on the way it is applied
// to V (row stride nn) and restarted. Returns -1 when a root does not converge (EIGS_FAILED).
//   per sweep: the search for a small sub-diagonal element and the search for two consecutive small ones are each a
//   set of independent tests, so every thread tests its own indices and the largest hit wins (atomic max) instead of
//   thread 0 walking down the diagonal; thread 0 then does the scalar branch work (roots, shifts) and all threads run
//   the bulge chase: row modification with one thread per column, barrier, column modification with one thread per
//   row, barrier.
template <class Cta>
SSTC_HD inline int eigs_schur(Cta &cta, double *H, int ld, double *V, double *val, int nn, EigsShared *sh,
        double *norm_out, EigsRot *log, int log_cap, EigsRot *chunk, int chunk_cap) {
    const int tid = cta.tid, nt = cta.nthreads;
    const double eps = eigs_eps();
    int log_n = 0;  // uniform: every thread counts the recorded steps
    // thread 0's private state
    int iter = 0;
    double exshift = 0.0;
    if (tid == 0) {
        double norm = 0.0;
        for (int i = 0; i < nn; i++) {
            for (int j = (i - 1 > 0 ? i - 1 : 0); j < nn; j++) {
                norm = norm + fabs(H[i * ld + j]);
            }
        }
        *norm_out = norm;
        sh->norm = norm;
        sh->n = nn - 1;
        sh->l = 0;
    }
    cta.sync();
    const double norm = sh->norm;
    for (;;) {
        const int n = sh->n;
        if (n < 0) {
            return log_n;
        }
        // look for a single small sub-diagonal element: the largest l in [1, n] that passes, else 0
        for (int l = 1 + tid; l <= n; l += nt) {
            double s = fabs(H[(l - 1) * ld + (l - 1)]) + fabs(H[l * ld + l]);
            if (s == 0.0) {
                s = norm;
            }
            if (fabs(H[l * ld + (l - 1)]) < eps * s) {
                cta.atomic_max(&sh->l, l);
            }
        }
        cta.sync();
        if (tid == 0) {
            const int l = sh->l;
            int mode = EIGS_NEXT;
            double s, w, x, y, z, p, q, r;
            if (l == n) {  // one root found
                H[n * ld + n] = H[n * ld + n] + exshift;
                val[2 * n] = H[n * ld + n];
                val[2 * n + 1] = 0.0;
                sh->n = n - 1;
                iter = 0;
            } else if (l == n - 1) {  // two roots found
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
                    p = p / r;
                    q = q / r;
                    sh->p = p;
                    sh->q = q;
                    sh->m = n;  // the pair's upper index
                    EigsRot &e = log[log_n];
                    e.x = p;
                    e.y = q;
                    e.z = e.q = e.r = 0.0;
                    e.k = n - 1;
                    e.kind = EIGS_ROT_PLANE;
                    mode = EIGS_REAL_PAIR;
                } else {  // complex pair
                    val[2 * (n - 1)] = x + p;
                    val[2 * n] = x + p;
                    val[2 * (n - 1) + 1] = z;
                    val[2 * n + 1] = -z;
                }
                sh->n = n - 2;
                iter = 0;
            } else {  // no convergence yet: form the shift
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
                } else {
                    sh->x = x;
                    sh->y = y;
                    sh->w = w;
                    sh->m = l;  // floor of the search below
                    mode = EIGS_QR_SWEEP;
                }
            }
            if (mode != EIGS_QR_SWEEP) {
                sh->l = 0;  // nobody else reads it in these modes; ready for the next search
            }
            sh->mode = mode;
        }
        cta.sync();
        const int mode = sh->mode;
        if (mode == EIGS_FAILED) {
            return -1;
        }
        if (mode == EIGS_NEXT) {
            continue;
        }
        if (mode == EIGS_REAL_PAIR) {
            const int n2 = sh->m;
            const double p = sh->p, q = sh->q;
            for (int j = n2 - 1 + tid; j < nn; j += nt) {  // row modification
                const double z = H[(n2 - 1) * ld + j];
                H[(n2 - 1) * ld + j] = q * z + p * H[n2 * ld + j];
                H[n2 * ld + j] = q * H[n2 * ld + j] - p * z;
            }
            cta.sync();
            for (int i = tid; i <= n2; i += nt) {  // column modification
                const double z = H[i * ld + (n2 - 1)];
                H[i * ld + (n2 - 1)] = q * z + p * H[i * ld + n2];
                H[i * ld + n2] = q * H[i * ld + n2] - p * z;
            }
            log_n++;  // thread 0 recorded the rotation before the barrier above
            cta.sync();
        } else {
            // look for two consecutive small sub-diagonal elements: the largest m in (l, n-2] that passes, else l
            const int l = sh->l;
            const double sx = sh->x, sy = sh->y, sw = sh->w;
            for (int m = l + 1 + tid; m <= n - 2; m += nt) {
                const double z = H[m * ld + m];
                double r = sx - z;
                double s = sy - z;
                double p = (r * s - sw) / H[(m + 1) * ld + m] + H[m * ld + (m + 1)];
                double q = H[(m + 1) * ld + (m + 1)] - z - r - s;
                r = H[(m + 2) * ld + (m + 1)];
                s = fabs(p) + fabs(q) + fabs(r);
                p = p / s;
                q = q / s;
                r = r / s;
                if (fabs(H[m * ld + (m - 1)]) * (fabs(q) + fabs(r)) <
                        eps * (fabs(p) * (fabs(H[(m - 1) * ld + (m - 1)]) + fabs(z) + fabs(H[(m + 1) * ld + (m + 1)])))) {
                    cta.atomic_max(&sh->m, m);
                }
            }
            cta.sync();
            const int m = sh->m;
            if (tid == 0) {
                sh->l = 0;  // everybody read it before the barrier above; the next search is barriers away
            }
            double p, q, r, x = sx, y, z, s;
            {
                z = H[m * ld + m];
                r = sx - z;
                s = sy - z;
                p = (r * s - sw) / H[(m + 1) * ld + m] + H[m * ld + (m + 1)];
                q = H[(m + 1) * ld + (m + 1)] - z - r - s;
                r = H[(m + 2) * ld + (m + 1)];
                s = fabs(p) + fabs(q) + fabs(r);
                p = p / s;
                q = q / s;
                r = r / s;
            }
            for (int i = m + 2 + tid; i <= n; i += nt) {
                H[i * ld + (i - 2)] = 0.0;
                if (i > m + 2) {
                    H[i * ld + (i - 3)] = 0.0;
                }
            }
            cta.sync();
            // double QR step involving rows l:n and columns m:n; every thread tracks the scalars redundantly
            for (int k = m; k <= n - 1; k++) {
                const bool not_last = (k != n - 1);
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
                        EigsRot &e = log[log_n];  // what this step does to V
                        e.x = x;
                        e.y = y;
                        e.z = z;
                        e.q = q;
                        e.r = r;
                        e.k = k;
                        e.kind = not_last ? EIGS_ROT_QR3 : EIGS_ROT_QR2;
                    }
                    const int ilast = n < k + 3 ? n : k + 3;
                    for (int i = tid; i <= ilast; i += nt) {  // column modification
                        double t = x * H[i * ld + k] + y * H[i * ld + (k + 1)];
                        if (not_last) {
                            t = t + z * H[i * ld + (k + 2)];
                            H[i * ld + (k + 2)] = H[i * ld + (k + 2)] - t * r;
                        }
                        H[i * ld + k] = H[i * ld + k] - t;
                        H[i * ld + (k + 1)] = H[i * ld + (k + 1)] - t * q;
                    }
                    log_n++;
                    cta.sync();
                    if (log_n == log_cap) {  // uniform
                        eigs_apply_rots(cta, V, nn, nn, log, log_n, chunk, chunk_cap);
                        log_n = 0;
                    }
                }
            }
            cta.sync();  // nobody still reads *sh when thread 0 starts its next scalar phase
        }
        if (log_n == log_cap) {  // a plane rotation filled the log
            eigs_apply_rots(cta, V, nn, nn, log, log_n, chunk, chunk_cap);
            log_n = 0;
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
