/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the FftService path of carsomyr/Shared.
 *
 * This file is a plain-C restatement of the reference's pure-Java FFT provider. It is the checker that the
 * parity tests, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs compare against
 * (or time). Nothing under shared_b200/ may call, link or load it: the product path is CUDA-only.
 *
 * Parity is PINNED: tests/test_oracle_fft.py replays every literal known-answer vector of the reference's
 * own FFT tests (src_test/org/shared/test/fft/ArrayFftTest.java:61-224,
 * src_test/org/shared/test/fft/ConvolutionCacheTest.java:56-91) through this code, and cross-checks it
 * against numpy's pocketfft.
 *
 * Restated functions (all paths relative to /root/reference):
 *   sst_oracle_factors    <- src/org/shared/fft/FftOps.java:111-150   (createFactors: even trial divisors only)
 *   sst_oracle_twiddles   <- src/org/shared/fft/FftOps.java:155-168   (createTwiddles)
 *   oracle_fft_rec        <- src/org/shared/fft/FftOps.java:173-206   (recursive decimation in time)
 *   oracle_fft_butterfly  <- src/org/shared/fft/FftOps.java:211-248   (generic radix-p butterfly)
 *   sst_oracle_fft        <- src/org/shared/fft/FftOps.java:55-106    (per-dimension driver, 1/N on inverse)
 *   sst_oracle_rfft       <- src/org/shared/fft/JavaFftService.java:53-64,192-207  (promote, c2c, keep half)
 *   sst_oracle_rifft      <- src/org/shared/fft/JavaFftService.java:67-84,115-183  (Hermitian completion, c2c^-1, Re)
 *
 * Differences from the Java, none numerical: offsets are 64-bit (the Java `len << 1` overflows int at 2^30
 * points), and arrays are malloc'd. Floating-point operation order inside each butterfly is kept exactly.
 * The Java runs on java.lang.Math.cos/sin; this runs on libm cos/sin (both are <= 1 ulp).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* FftOps.java:111-150. Writes [p0, m0, p1, m1, ...] into `out` (capacity >= 128) and returns the count. */
int sst_oracle_factors(int num, int *out) {
    int n = 0;
    int p = 2;
    int upper = (int) floor(sqrt((double) num));
    do {
        for (; num % p != 0;) {
            /* FftOps.java:122-131 -- the `case 1` arm is dead code; p always advances by two from 2. */
            p += 2;
            p = p > upper ? num : p;
        }
        num /= p;
        out[n++] = p;
        out[n++] = num;
    } while (num > 1);
    return n;
}

/* FftOps.java:155-168. `res` holds 2*n doubles. */
void sst_oracle_twiddles(int n, int direction, double *res) {
    for (int i = 0; i < n; i++) {
        /* Java: (-2 * direction * Math.PI * i) / n, evaluated left to right in double. */
        double phase = (-2 * direction * M_PI * i) / n;
        res[2 * i] = cos(phase);
        res[2 * i + 1] = sin(phase);
    }
}

/* FftOps.java:211-248. */
static void oracle_fft_butterfly(double *dst, int64_t dstOffsetCurrent, double *scratch, int p, int n,
        const double *twiddles, int size, int strideCurrent) {

    int64_t dstOffsetIncr = (int64_t) n << 1;
    int64_t twiddleOffsetIncr = ((int64_t) strideCurrent * n) << 1;
    int64_t modulus = (int64_t) size << 1;

    for (int i = 0; i < n; i++) {

        int64_t dstOffset = ((int64_t) i << 1) + dstOffsetCurrent;
        for (int j = 0; j < p; j++, dstOffset += dstOffsetIncr) {
            scratch[2 * j] = dst[dstOffset];
            scratch[2 * j + 1] = dst[dstOffset + 1];
        }

        dstOffset = ((int64_t) i << 1) + dstOffsetCurrent;
        int64_t twiddleOffset = ((int64_t) strideCurrent * i) << 1;

        for (int j = 0; j < p; j++, dstOffset += dstOffsetIncr, twiddleOffset += twiddleOffsetIncr) {

            dst[dstOffset] = scratch[0];
            dst[dstOffset + 1] = scratch[1];

            int64_t modOffset = twiddleOffset;
            for (int k = 1; k < p; k++, modOffset = (modOffset + twiddleOffset) % modulus) {
                int so = 2 * k;
                dst[dstOffset] += scratch[so] * twiddles[modOffset] - scratch[so + 1] * twiddles[modOffset + 1];
                dst[dstOffset + 1] += scratch[so] * twiddles[modOffset + 1] + scratch[so + 1] * twiddles[modOffset];
            }
        }
    }
}

/* FftOps.java:173-206. */
static void oracle_fft_rec(const double *src, int64_t srcOffset, double *dst, int64_t dstOffset, double *scratch,
        const int *factors, int factorIndexCurrent, const double *twiddles, int size, int64_t stride,
        int strideCurrent) {

    int p = factors[factorIndexCurrent];
    int m = factors[factorIndexCurrent + 1];

    int64_t srcOffsetIncr = ((int64_t) strideCurrent * stride) << 1;
    int64_t dstOffsetIncr = (int64_t) m << 1;
    int64_t upper = dstOffset + (((int64_t) p * m) << 1);

    if (m == 1) {
        int64_t s = srcOffset;
        for (int64_t d = dstOffset; d < upper; s += srcOffsetIncr, d += 2) {
            dst[d] = src[s];
            dst[d + 1] = src[s + 1];
        }
    } else {
        int64_t s = srcOffset;
        for (int64_t d = dstOffset; d < upper; s += srcOffsetIncr, d += dstOffsetIncr) {
            oracle_fft_rec(src, s, dst, d, scratch, factors, factorIndexCurrent + 2, twiddles, size, stride,
                    strideCurrent * p);
        }
    }

    oracle_fft_butterfly(dst, dstOffset, scratch, p, m, twiddles, size, strideCurrent);
}

static int64_t product(int ndims, const int *dims) {
    int64_t len = 1;
    for (int i = 0; i < ndims; i++) {
        len *= dims[i];
    }
    return len;
}

/*
 * FftOps.java:55-106. direction = +1 forward, -1 inverse (scaled by 1/N). `in` and `out` hold 2*N doubles
 * (interleaved re/im, row-major); they may not alias. Returns 0, or -1 on bad arguments / allocation failure.
 */
int sst_oracle_fft(int direction, int ndims, const int *dims, const double *in, double *out) {

    if (ndims <= 0 || (direction != 1 && direction != -1)) {
        return -1;
    }
    for (int i = 0; i < ndims; i++) {
        if (dims[i] <= 0) {
            return -1;
        }
    }

    int64_t len = product(ndims, dims);
    int maxDim = 0;
    for (int i = 0; i < ndims; i++) {
        maxDim = dims[i] > maxDim ? dims[i] : maxDim;
    }

    double *outTmp = (double *) malloc(sizeof(double) * 2 * (size_t) len);
    double *scratch = (double *) malloc(sizeof(double) * 2 * (size_t) maxDim);
    double *twiddles = (double *) malloc(sizeof(double) * 2 * (size_t) maxDim);
    if (!outTmp || !scratch || !twiddles) {
        free(outTmp);
        free(scratch);
        free(twiddles);
        return -1;
    }

    memcpy(out, in, sizeof(double) * 2 * (size_t) len);

    for (int dim = 0; dim < ndims; dim++) {

        int dimSize = dims[dim];
        int64_t dimStride = len / dimSize;
        int factors[128];

        sst_oracle_factors(dimSize, factors);
        sst_oracle_twiddles(dimSize, direction, twiddles);

        int64_t srcOffset = 0, dstOffset = 0;
        for (int64_t j = 0; j < dimStride; j++, srcOffset += 2, dstOffset += (int64_t) dimSize << 1) {
            oracle_fft_rec(out, srcOffset, outTmp, dstOffset, scratch, factors, 0, twiddles, dimSize, dimStride, 1);
        }

        memcpy(out, outTmp, sizeof(double) * 2 * (size_t) len);
    }

    if (direction == -1) {
        double factor = 1.0 / (double) len;
        for (int64_t i = 0; i < 2 * len; i++) {
            out[i] *= factor;
        }
    }

    free(outTmp);
    free(scratch);
    free(twiddles);
    return 0;
}

/*
 * JavaFftService.java:53-64 + fullToReduced :192-207. `dims` are the REAL logical dims; `in` has N doubles,
 * `out` has 2 * (N / last) * (last / 2 + 1) doubles.
 */
int sst_oracle_rfft(int ndims, const int *dims, const double *in, double *out) {

    if (ndims <= 0) {
        return -1;
    }
    int64_t len = product(ndims, dims);
    int last = dims[ndims - 1];
    int half = last / 2 + 1;
    int64_t rows = len / last;

    double *cin = (double *) calloc(2 * (size_t) len, sizeof(double));
    double *cout = (double *) malloc(sizeof(double) * 2 * (size_t) len);
    if (!cin || !cout) {
        free(cin);
        free(cout);
        return -1;
    }
    for (int64_t i = 0; i < len; i++) {
        cin[2 * i] = in[i]; /* tocRe(): ElementOps.java R_TO_C_RE */
    }
    int rc = sst_oracle_fft(+1, ndims, dims, cin, cout);
    if (rc == 0) {
        for (int64_t r = 0; r < rows; r++) {
            memcpy(out + 2 * r * half, cout + 2 * r * last, sizeof(double) * 2 * (size_t) half);
        }
    }
    free(cin);
    free(cout);
    return rc;
}

/*
 * JavaFftService.java:67-84 + reducedToFull :115-183. The Java builds the full spectrum as: copy the reduced
 * half (:136); for k_last >= half take conj(reduced[(-k_0) mod d_0, ..., (-k_{n-2}) mod d_{n-2}, d_last - k_last])
 * (the per-dimension mirrored splices :155-174, the conjugation and the shifted map :178-182). Then inverse
 * c2c and real part (torRe). `in` has 2 * rows * half doubles, `out` has N doubles.
 */
int sst_oracle_rifft(int ndims, const int *dims, const double *in, double *out) {

    if (ndims <= 0) {
        return -1;
    }
    int64_t len = product(ndims, dims);
    int last = dims[ndims - 1];
    int half = last / 2 + 1;
    int64_t rows = len / last;

    double *full = (double *) malloc(sizeof(double) * 2 * (size_t) len);
    double *cout = (double *) malloc(sizeof(double) * 2 * (size_t) len);
    if (!full || !cout) {
        free(full);
        free(cout);
        return -1;
    }

    for (int64_t r = 0; r < rows; r++) {
        /* mirrored row index: negate every leading logical index modulo its dimension */
        int64_t rem = r, mr = 0, mult = 1;
        for (int d = ndims - 2; d >= 0; d--) {
            int64_t k = rem % dims[d];
            rem /= dims[d];
            mr += ((dims[d] - k) % dims[d]) * mult;
            mult *= dims[d];
        }
        for (int k = 0; k < half; k++) {
            full[2 * (r * last + k)] = in[2 * (r * half + k)];
            full[2 * (r * last + k) + 1] = in[2 * (r * half + k) + 1];
        }
        for (int k = half; k < last; k++) {
            full[2 * (r * last + k)] = in[2 * (mr * half + (last - k))];
            full[2 * (r * last + k) + 1] = -in[2 * (mr * half + (last - k)) + 1];
        }
    }

    int rc = sst_oracle_fft(-1, ndims, dims, full, cout);
    if (rc == 0) {
        for (int64_t i = 0; i < len; i++) {
            out[i] = cout[2 * i]; /* torRe(): C_TO_R_RE */
        }
    }
    free(full);
    free(cout);
    return rc;
}
