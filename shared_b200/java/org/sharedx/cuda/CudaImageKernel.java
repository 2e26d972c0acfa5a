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

public final class CudaImageKernel implements ImageKernel {

    /**
     * Public and without arguments: {@code Services.createService} instantiates providers reflectively. Refuses to exist
     * unless libsst_cuda's {@code JNI_OnLoad} found a CUDA device and registered the providers.
     */
    public CudaImageKernel() {
        if (!Library.isInitialized()) {
            throw new RuntimeException("libsst_cuda is not loaded or found no sm_100 device: "
                    + "the CUDA ImageKernel cannot be instantiated (there is no CPU fallback)");
        }
    }

    /** sstc_integral_image (JNI export Java_org_sharedx_cuda_CudaImageKernel_createIntegralImage). */
    @Override
    public native void createIntegralImage(double[] srcV, int[] srcD, int[] srcS, double[] dstV, int[] dstD, int[] dstS);

    /** sstc_integral_histogram (JNI export Java_org_sharedx_cuda_CudaImageKernel_createIntegralHistogram). */
    @Override
    public native void createIntegralHistogram(double[] srcV, int[] srcD, int[] srcS, int[] memV, double[] dstV, int[] dstD,
            int[] dstS);
}
