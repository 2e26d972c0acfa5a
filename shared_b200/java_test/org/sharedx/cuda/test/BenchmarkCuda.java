/**
 * The reference's benchmark ({@code src_test/org/sharedx/test/BenchmarkSpecification.java:43-53}: 2048 repetitions of
 * [complex eMul, 2-D inverse FFT with the 1/N scale] on a 256 x 256 complex array, after 64 fft / ifft warm-ups) under the CUDA
 * providers, twice:
 * <ul>
 * <li>{@link #testConvolve()} is {@code BenchmarkJava.testConvolve} (BenchmarkJava.java:92-104) verbatim in structure: every
 * operation crosses the provider boundary with JVM arrays, so each repetition pays two H2D and two D2H copies of 1 MiB.</li>
 * <li>{@link #testConvolveResident()} keeps the arrays in HBM ({@link CudaComplexArray}): the counterpart of
 * {@code BenchmarkNative} / {@code Benchmark.cpp:59-116}, whose loop never leaves native memory. bench.py's
 * {@code extra_configs} entry "C1b" times exactly this loop through the same C ABI.</li>
 * </ul>
 */
package org.sharedx.cuda.test;

import static org.shared.array.ArrayBase.fftService;

import org.junit.BeforeClass;
import org.junit.Test;
import org.shared.array.ComplexArray;
import org.shared.fft.ConvolutionCache;
import org.sharedx.cuda.CudaComplexArray;
import org.sharedx.test.BenchmarkSpecification;

public class BenchmarkCuda implements BenchmarkSpecification {

    public BenchmarkCuda() {
    }

    /** Plan creation, twiddle tables and staging buffers happen here, outside the timed tests. */
    @BeforeClass
    final public static void initClass() {

        fftService.setHint("mode", "measure"); // accepted and ignored: plans are deterministic

        ComplexArray tmp = new ComplexArray(SIZE, SIZE, 2);

        for (int i = 0; i < 64; i++) {
            tmp.fft().ifft();
        }
    }

    @Test
    @Override
    public void testConvolve() {

        ComplexArray kernel = new ComplexArray(SIZE, SIZE, 2);
        ComplexArray im = new ComplexArray(SIZE, SIZE, 2);

        ConvolutionCache cc = ConvolutionCache.getInstance();

        for (int i = 0; i < N_REPS; i++) {
            im.eMul(cc.get(kernel, im.dims())).ifft();
        }
    }

    @Test
    public void testConvolveResident() {

        CudaComplexArray kernel = new CudaComplexArray(SIZE, SIZE, 2);
        CudaComplexArray im = new CudaComplexArray(SIZE, SIZE, 2);
        CudaComplexArray tmp = new CudaComplexArray(SIZE, SIZE, 2);

        try {

            for (int i = 0; i < N_REPS; i++) {
                im.eMul(kernel, tmp).ifft(tmp); // product into tmp, inverse transform in place on the device
            }

            CudaComplexArray.synchronize();

        } finally {

            kernel.close();
            im.close();
            tmp.close();
        }
    }
}
