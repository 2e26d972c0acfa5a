/**
 * The B200 (sm_100a) provider of {@link org.shared.array.kernel.ArrayKernel}: a drop-in for
 * {@code JavaArrayKernel} and {@code org.shared.array.jni.NativeArrayKernel} behind
 * {@code org.shared.array.kernel.ModalArrayKernel}. Callers change only the service binding:
 *
 * <pre>
 * ((RegistryClassLoader) Thread.currentThread().getContextClassLoader()).loadLibrary("lib/sst_cuda");
 * ArrayBase.opKernel.useRegisteredKernel();
 * </pre>
 *
 * Every member of the SPI is a native method served by libsst_cuda.so: the JNI glue
 * (shared_b200/csrc/jni/jni_glue.cpp) has one export per method, each of which pins the arrays, calls ONE host-tier
 * function of the C ABI (include/sst_cuda.h, named in the comment of each method below), releases and rethrows the
 * library's error text as a {@link RuntimeException}. All arithmetic runs on the GPU; there is no CPU fallback.
 * The parameter lists are those of the interface (ArrayKernel.java:285-674).
 */
package org.sharedx.cuda;

import org.shared.array.kernel.ArrayKernel;
import org.shared.array.sparse.SparseArrayState;
import org.shared.metaclass.Library;

public final class CudaArrayKernel implements ArrayKernel {

    /**
     * Public and without arguments because {@code Services.createService} instantiates providers reflectively
     * (Services.java:72-119). Refuses to exist unless libsst_cuda's {@code JNI_OnLoad} found a CUDA device and registered
     * the providers ({@code Library.initialized}, set by the glue like Library.cpp:77 does).
     */
    public CudaArrayKernel() {
        if (!Library.isInitialized()) {
            throw new RuntimeException("libsst_cuda is not loaded or found no sm_100 device: "
                    + "the CUDA ArrayKernel cannot be instantiated (there is no CPU fallback)");
        }
    }

    // ---- generator state: sstc_srand --------------------------------------------------------------------------------

    @Override
    public native void randomize();

    @Override
    public native void derandomize();

    // ---- data movement: sstc_map / sstc_slice (Object[] arrays are moved by the glue on the JVM side) ------------------

    @Override
    public native void map(int[] bounds, Object srcV, int[] srcD, int[] srcS, Object dstV, int[] dstD, int[] dstS);

    @Override
    public native void slice(int[] slices, Object srcV, int[] srcD, int[] srcS, Object dstV, int[] dstD, int[] dstS);

    // ---- along dimensions: sstc_rrop (reduce), sstc_riop (positions / sort), sstc_rdop (inclusive scan) ----------------

    @Override
    public native void rrOp(int type, double[] srcV, int[] srcD, int[] srcS, double[] dstV, int[] dstD, int[] dstS,
            int... opDims);

    @Override
    public native void riOp(int type, double[] srcV, int[] srcD, int[] srcS, int[] dstV, int dim);

    @Override
    public native void rdOp(int type, double[] srcV, int[] srcD, int[] srcS, double[] dstV, int... opDims);

    // ---- whole-array folds: sstc_raop -> double, sstc_caop -> {re, im} ---------------------------------------------------

    @Override
    public native double raOp(int type, double[] srcV);

    @Override
    public native double[] caOp(int type, double[] srcV);

    // ---- element-wise, in place: sstc_ruop / sstc_cuop / sstc_iuop -------------------------------------------------------

    @Override
    public native void ruOp(int type, double a, double[] srcV);

    @Override
    public native void cuOp(int type, double aRe, double aIm, double[] srcV);

    @Override
    public native void iuOp(int type, int a, int[] srcV);

    // ---- element-wise, binary and conversions: sstc_eop / sstc_convert (runtime type dispatch in the glue) -----------------

    @Override
    public native void eOp(int type, Object lhsV, Object rhsV, Object dstV, boolean complex);

    @Override
    public native void convert(int type, Object srcV, boolean isSrcComplex, Object dstV, boolean isDstComplex);

    // ---- matrices: sstc_mul (FP64 tensor-pipe GEMM), sstc_diag -----------------------------------------------------------

    @Override
    public native void mul(double[] lhsV, double[] rhsV, int lr, int rc, double[] dstV, boolean complex);

    @Override
    public native void diag(double[] srcV, double[] dstV, int size, boolean complex);

    // ---- decompositions --------------------------------------------------------------------------------------------------

    /** sstc_svd: one-sided Jacobi on the GPU; singular vectors agree with JAMA's up to the sign of each pair. */
    @Override
    public native void svd(double[] srcV, int srcStrideRow, int srcStrideCol, double[] uV, double[] sV, double[] vV,
            int nRows, int nCols);

    /**
     * sstc_eigs: Hessenberg reduction + double-shift QR in one CTA, back-substitution with one thread per eigenvalue;
     * bit-identical to the reference's native kernel.
     */
    @Override
    public native void eigs(double[] srcV, double[] vecV, double[] valV, int size);

    /** sstc_invert: LU with partial pivoting (JAMA's pivot rule) and the two triangular solves on the GPU. */
    @Override
    public native void invert(double[] srcV, double[] dstV, int size);

    // ---- index extraction and sparse arrays ------------------------------------------------------------------------------

    /** sstc_find; the {@code int[]} result is allocated by the glue. */
    @Override
    public native int[] find(int[] srcV, int[] srcD, int[] srcS, int[] logical);

    /**
     * sstc_insert_sparse: pair sort, scan-ranked merge and per-dimension metadata on the GPU. The glue constructs the
     * {@link SparseArrayState}; {@code Object[]} values are gathered on the JVM side from positions the device computes.
     */
    @Override
    public native <V> SparseArrayState<V> insertSparse(V oldV, int[] oldD, int[] oldS, int[] oldDo, int[] oldI, V newV,
            int[] newI);

    /** sstc_slice_sparse. */
    @Override
    public native <V> SparseArrayState<V> sliceSparse(int[] slices, V srcV, int[] srcD, int[] srcS, int[] srcDo, int[] srcI,
            int[] srcIo, int[] srcIi, V dstV, int[] dstD, int[] dstS, int[] dstDo, int[] dstI, int[] dstIo, int[] dstIi);
}
