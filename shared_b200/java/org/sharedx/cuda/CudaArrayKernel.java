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
 * The sixteen operations on the FFT / array-kernel hot path are native (JNI glue:
 * shared_b200/csrc/jni/jni_glue.cpp, one export per method, same signatures as NativeArrayKernel.java:56-154).
 * The interface members outside that path (invert, svd, eigs, find, insertSparse, sliceSparse) are native as well:
 * every member of the SPI runs on the GPU, and this provider has no CPU fallback.
 */
package org.sharedx.cuda;

import org.shared.array.kernel.ArrayKernel;
import org.shared.array.sparse.SparseArrayState;
import org.shared.metaclass.Library;
import org.shared.util.Control;

public class CudaArrayKernel implements ArrayKernel {

    /**
     * Default constructor. Checks the validity of native bindings.
     */
    public CudaArrayKernel() {
        Control.checkTrue(Library.isInitialized(), //
                "Could not instantiate native bindings -- Linking failed");
    }

    @Override
    final public native void randomize();

    @Override
    final public native void derandomize();

    @Override
    final public native void map(int[] bounds, //
            Object srcV, int[] srcD, int[] srcS, //
            Object dstV, int[] dstD, int[] dstS);

    @Override
    final public native void slice(int[] slices, //
            Object srcV, int[] srcD, int[] srcS, //
            Object dstV, int[] dstD, int[] dstS);

    @Override
    final public native void rrOp(int type, //
            double[] srcV, int[] srcD, int[] srcS, //
            double[] dstV, int[] dstD, int[] dstS, //
            int... opDims);

    @Override
    final public native void riOp(int type, //
            double[] srcV, int[] srcD, int[] srcS, int[] dstV, int dim);

    @Override
    final public native void rdOp(int type, //
            double[] srcV, int[] srcD, int[] srcS, double[] dstV, //
            int... opDims);

    @Override
    final public native double raOp(int type, double[] srcV);

    @Override
    final public native double[] caOp(int type, double[] srcV);

    @Override
    final public native void ruOp(int type, double a, double[] srcV);

    @Override
    final public native void cuOp(int type, double aRe, double aIm, double[] srcV);

    @Override
    final public native void iuOp(int type, int a, int[] srcV);

    @Override
    final public native void eOp(int type, Object lhsV, Object rhsV, Object dstV, boolean complex);

    @Override
    final public native void convert(int type, //
            Object srcV, boolean isSrcComplex, //
            Object dstV, boolean isDstComplex);

    @Override
    final public native void mul(double[] lhsV, double[] rhsV, int lr, int rc, double[] dstV, boolean complex);

    @Override
    final public native void diag(double[] srcV, double[] dstV, int size, boolean complex);

    //

    /** JNI: Java_org_sharedx_cuda_CudaArrayKernel_svd -> sstc_svd (one-sided Jacobi on the GPU). */
    @Override
    final public native void svd(double[] srcV, int srcStrideRow, int srcStrideCol, //
            double[] uV, double[] sV, double[] vV, int nRows, int nCols);

    /**
     * JNI: Java_org_sharedx_cuda_CudaArrayKernel_eigs -> sstc_eigs (Hessenberg + double-shift QR in one CTA,
     * back-substitution with one thread per eigenvalue; bit-identical to the reference's native kernel).
     */
    @Override
    final public native void eigs(double[] srcV, double[] vecV, double[] valV, int size);

    /** JNI: Java_org_sharedx_cuda_CudaArrayKernel_invert -> sstc_invert (LU with partial pivoting on the GPU). */
    @Override
    final public native void invert(double[] srcV, double[] dstV, int size);

    /** JNI: Java_org_sharedx_cuda_CudaArrayKernel_find -> sstc_find; the result array is allocated by the glue. */
    @Override
    final public native int[] find(int[] srcV, int[] srcD, int[] srcS, int[] logical);

    /**
     * JNI: Java_org_sharedx_cuda_CudaArrayKernel_insertSparse -> sstc_insert_sparse (pair sort, scan-ranked merge and
     * per-dimension metadata on the GPU); the SparseArrayState is constructed by the glue. Object[] values are gathered
     * on the JVM side from the positions the device computes.
     */
    @Override
    final public native <V> SparseArrayState<V> insertSparse( //
            V oldV, int[] oldD, int[] oldS, int[] oldDo, int[] oldI, //
            V newV, int[] newI);

    /** JNI: Java_org_sharedx_cuda_CudaArrayKernel_sliceSparse -> sstc_slice_sparse. */
    @Override
    final public native <V> SparseArrayState<V> sliceSparse(int[] slices, //
            V srcV, int[] srcD, int[] srcS, int[] srcDo, //
            int[] srcI, int[] srcIo, int[] srcIi, //
            V dstV, int[] dstD, int[] dstS, int[] dstDo, //
            int[] dstI, int[] dstIo, int[] dstIi);
}
