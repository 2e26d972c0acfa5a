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

    /** Hints this provider knows, with their current values (FFTW's planner hints are accepted for compatibility). */
    final private java.util.Map<String, String> hints = new java.util.concurrent.ConcurrentHashMap<String, String>();

    /** Values "mode" may take (FFTW's planner rigor levels; plans here are deterministic, so they change nothing). */
    final private static java.util.Set<String> MODES = new java.util.HashSet<String>(
            java.util.Arrays.asList("estimate", "measure", "patient", "exhaustive"));

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

    /** JNI: Java_org_sharedx_cuda_CudaFftService_setDevices -> sstc_fft_set_devices. */
    final native void setDevices(int[] devices);

    /**
     * Hints: {@code "mode"} and {@code "wisdom"} as {@code FftwService} takes them; {@code "devices"}, a comma-separated
     * list of CUDA device ordinals -- with two or more, 3-D complex transforms whose shape allows it run slab-sharded
     * over all of them, every device moving its part of the array over its own PCIe link.
     */
    @Override
    public void setHint(String name, String value) {

        if ("mode".equals(name)) {

            if (!MODES.contains(value)) {
                throw new IllegalArgumentException("Invalid execution mode");
            }

        } else if ("devices".equals(name)) {

            String[] parts = value.trim().isEmpty() ? new String[0] : value.split(",");
            int[] devices = new int[parts.length];

            for (int i = 0; i < parts.length; i++) {

                try {
                    devices[i] = Integer.parseInt(parts[i].trim());
                } catch (NumberFormatException e) {
                    throw new IllegalArgumentException("Invalid device list");
                }
            }

            setDevices(devices);

        } else if (!"wisdom".equals(name)) {

            throw new IllegalArgumentException("Unknown hint");
        }

        this.hints.put(name, "wisdom".equals(name) ? "" : value);
    }

    @Override
    public String getHint(String name) {

        if (!"mode".equals(name) && !"wisdom".equals(name) && !"devices".equals(name)) {
            throw new IllegalArgumentException("Unknown hint");
        }

        String value = this.hints.get(name);
        return value != null ? value : ("mode".equals(name) ? "estimate" : "");
    }
}
