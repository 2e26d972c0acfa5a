/**
 * The B200 (sm_100a) provider of {@link org.shared.fft.FftService}: a drop-in for
 * {@code org.shared.fft.JavaFftService} and {@code org.sharedx.fftw.FftwService} behind
 * {@code org.shared.fft.ModalFftService}. Callers change only the service binding:
 *
 * <pre>
 * ((RegistryClassLoader) Thread.currentThread().getContextClassLoader()).loadLibrary("lib/sst_cuda");
 * ArrayBase.fftService.useRegisteredService();
 * </pre>
 *
 * All arithmetic happens in libsst_cuda.so (hand-written CUDA behind the C ABI of include/sst_cuda.h); this
 * class only forwards. There is no CPU fallback: without the library the constructor fails.
 */
package org.sharedx.cuda;

import org.shared.fft.FftService;
import org.shared.metaclass.Library;

public class CudaFftService implements FftService {

    /** Transform kinds; the values of {@code org.sharedx.fftw.Plan} (Plan.java:73-88) and SSTC_FFT_* . */
    final public static int R_TO_C = 0;
    final public static int C_TO_R = 1;
    final public static int FORWARD = 2;
    final public static int BACKWARD = 3;

    /** FFTW's planner modes are accepted for compatibility and ignored: plans here are deterministic. */
    volatile String mode = "estimate";

    /**
     * Public and without arguments: {@code Services.createService} instantiates providers reflectively. Refuses to exist
     * unless libsst_cuda's {@code JNI_OnLoad} found a CUDA device and registered the providers (the same gate the
     * reference's native providers have, NativeArrayKernel.java:46-49).
     */
    public CudaFftService() {
        if (!Library.isInitialized()) {
            throw new RuntimeException("libsst_cuda is not loaded or found no sm_100 device: "
                    + "the CUDA FftService cannot be instantiated (there is no CPU fallback)");
        }
    }

    /** JNI: Java_org_sharedx_cuda_CudaFftService_transform -> sstc_fft. */
    final native void transform(int kind, int[] dims, double[] in, double[] out);

    @Override
    public void rfft(int[] dims, double[] in, double[] out) {
        transform(R_TO_C, dims, in, out);
    }

    @Override
    public void rifft(int[] dims, double[] in, double[] out) {
        transform(C_TO_R, dims, in, out);
    }

    @Override
    public void fft(int[] dims, double[] in, double[] out) {
        transform(FORWARD, dims, in, out);
    }

    @Override
    public void ifft(int[] dims, double[] in, double[] out) {
        transform(BACKWARD, dims, in, out);
    }

    @Override
    public void setHint(String name, String value) {

        if (name.equals("mode")) {

            if (value.equals("estimate") || value.equals("measure") //
                    || value.equals("patient") || value.equals("exhaustive")) {

                this.mode = value;

            } else {

                throw new IllegalArgumentException("Invalid execution mode");
            }

        } else if (name.equals("wisdom")) {

            // Nothing to import: there is no planner state.

        } else {

            throw new IllegalArgumentException("Unknown hint");
        }
    }

    @Override
    public String getHint(String name) {

        if (name.equals("mode")) {

            return this.mode;

        } else if (name.equals("wisdom")) {

            return "";

        } else {

            throw new IllegalArgumentException("Unknown hint");
        }
    }
}
