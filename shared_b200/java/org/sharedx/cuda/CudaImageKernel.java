/**
 * The B200 (sm_100a) provider of {@link org.shared.image.kernel.ImageKernel}: a drop-in for
 * {@code org.shared.image.kernel.JavaImageKernel} and {@code org.shared.image.jni.NativeImageKernel} behind
 * {@code org.shared.image.kernel.ModalImageKernel} (ModalImageKernel.java:86-97). Callers change only the
 * service binding:
 *
 * <pre>
 * ((RegistryClassLoader) Thread.currentThread().getContextClassLoader()).loadLibrary("lib/sst_cuda");
 * ImageOps.imKernel.useRegisteredKernel();
 * </pre>
 *
 * Both methods forward to libsst_cuda.so (sstc_integral_image / sstc_integral_histogram, include/sst_cuda.h);
 * argument meaning, mutation of {@code dstV} and exception messages are those of NativeImageKernel
 * (native/src/shared/NativeImageKernel.cpp:32-247). There is no CPU fallback.
 */
package org.sharedx.cuda;

import org.shared.image.kernel.ImageKernel;
import org.shared.metaclass.Library;
import org.shared.util.Control;

public class CudaImageKernel implements ImageKernel {

    /**
     * Default constructor (instantiated reflectively by {@code Services.createService}); same linking check as
     * {@code NativeImageKernel} (NativeImageKernel.java:45-49).
     */
    public CudaImageKernel() {
        Control.checkTrue(Library.isInitialized(), //
                "Could not instantiate native bindings -- Linking failed");
    }

    /** JNI: Java_org_sharedx_cuda_CudaImageKernel_createIntegralImage -> sstc_integral_image. */
    @Override
    final public native void createIntegralImage( //
            double[] srcV, int[] srcD, int[] srcS, //
            double[] dstV, int[] dstD, int[] dstS);

    /** JNI: Java_org_sharedx_cuda_CudaImageKernel_createIntegralHistogram -> sstc_integral_histogram. */
    @Override
    final public native void createIntegralHistogram( //
            double[] srcV, int[] srcD, int[] srcS, int[] memV, //
            double[] dstV, int[] dstD, int[] dstS);
}
