/**
 * A complex array whose values live in HBM: the device-resident twin of {@code org.shared.array.ComplexArray} for callers
 * that chain operations (SURVEY.md 8 f-1). Every reference array is a JVM {@code double[]} ({@code ProtoArray.java:56}), so a
 * provider behind the SPIs crosses PCIe twice per call; this class holds a device pointer instead and forwards to the DEVICE
 * tier of the C ABI (include/sst_cuda.h: {@code sstc_eop_dev}, {@code sstc_fft_plan_execute}, {@code sstc_fftshift_dev} ..),
 * asynchronously on the library's stream. It mirrors shared_b200/device_array.py (the twin the parity tests drive).
 * <p>
 * Dims carry the trailing 2 of a {@code ComplexArray}; values are row-major interleaved re / im. Not a subclass of
 * {@code AbstractComplexArray}: that class owns a {@code double[]} by construction. {@link #toComplexArray()} and
 * {@link #CudaComplexArray(ComplexArray)} are the two copies in and out.
 * </p>
 * <p>
 * Source only in this repository's build image (no JDK); the natives are implemented in shared_b200/csrc/jni/jni_glue.cpp.
 * </p>
 */
package org.sharedx.cuda;

import java.io.Closeable;

import org.shared.array.ComplexArray;

// java.io.Closeable, not AutoCloseable: the reference builds at source level 1.6 (build.xml:8-9)
public class CudaComplexArray implements Closeable {

    final int[] dims;
    final int length; // doubles
    long devicePointer;

    /** A zero-filled array of the given dims (the last one must be 2). */
    public CudaComplexArray(int... dims) {

        if (dims.length < 2 || dims[dims.length - 1] != 2) {
            throw new IllegalArgumentException("Invalid dimensions");
        }

        long n = 1;

        for (int d : dims) {

            n *= d;

            if (d < 0 || n > Integer.MAX_VALUE) {
                throw new IllegalArgumentException("Invalid dimensions");
            }
        }

        this.dims = dims.clone();
        this.length = (int) n;
        this.devicePointer = allocate(this.length);
    }

    /** Uploads a host array (one H2D copy). */
    public CudaComplexArray(ComplexArray src) {
        this(src.dims());
        upload(this.devicePointer, src.values());
    }

    /** Downloads into a new host array (one D2H copy; synchronises). */
    public ComplexArray toComplexArray() {
        double[] values = new double[this.length];
        download(this.devicePointer, values);
        return new ComplexArray(values, this.dims);
    }

    public int[] dims() {
        return this.dims.clone();
    }

    /** {@code dst = this .* rhs} (ComplexArray.eMul, CE_MUL through {@code sstc_eop_dev}); returns {@code dst}. */
    public CudaComplexArray eMul(CudaComplexArray rhs, CudaComplexArray dst) {
        checkSameShape(rhs);
        checkSameShape(dst);
        eMul(this.devicePointer, rhs.devicePointer, dst.devicePointer, this.length);
        return dst;
    }

    /** Forward transform of this array into {@code dst} (which may be this array). */
    public CudaComplexArray fft(CudaComplexArray dst) {
        checkSameShape(dst);
        transform(CudaFftService.FORWARD, logicalDims(), this.devicePointer, dst.devicePointer);
        return dst;
    }

    /** Backward transform (scaled by 1/N) of this array into {@code dst} (which may be this array). */
    public CudaComplexArray ifft(CudaComplexArray dst) {
        checkSameShape(dst);
        transform(CudaFftService.BACKWARD, logicalDims(), this.devicePointer, dst.devicePointer);
        return dst;
    }

    /** {@code AbstractComplexArray.fftShift} (AbstractComplexArray.java:67-97) into {@code dst} (not this array). */
    public CudaComplexArray fftShift(CudaComplexArray dst) {
        checkSameShape(dst);
        fftShift(this.devicePointer, dst.devicePointer, this.dims, false);
        return dst;
    }

    public CudaComplexArray ifftShift(CudaComplexArray dst) {
        checkSameShape(dst);
        fftShift(this.devicePointer, dst.devicePointer, this.dims, true);
        return dst;
    }

    /** Waits for everything issued so far (and raises what a failed asynchronous launch left behind). */
    public static native void synchronize();

    @Override
    public void close() {

        if (this.devicePointer != 0) {
            free(this.devicePointer);
            this.devicePointer = 0;
        }
    }

    int[] logicalDims() {
        return java.util.Arrays.copyOf(this.dims, this.dims.length - 1);
    }

    void checkSameShape(CudaComplexArray other) {

        if (!java.util.Arrays.equals(this.dims, other.dims) || other.devicePointer == 0 || this.devicePointer == 0) {
            throw new IllegalArgumentException("Dimension mismatch");
        }
    }

    // JNI: Java_org_sharedx_cuda_CudaComplexArray_* -> sstc_malloc / sstc_free / sstc_h2d / sstc_d2h / sstc_eop_dev /
    // sstc_fft_plan_execute (plans cached per kind and dims) / sstc_fftshift_dev / sstc_stream_sync

    static native long allocate(int nDoubles);

    static native void free(long pointer);

    static native void upload(long pointer, double[] values);

    static native void download(long pointer, double[] values);

    static native void eMul(long lhs, long rhs, long dst, int nDoubles);

    static native void transform(int kind, int[] dims, long in, long out);

    static native void fftShift(long src, long dst, int[] dims, boolean inverse);
}
