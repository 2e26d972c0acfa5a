/**
 * The acceptance driver of the CUDA providers: the reference's own JUnit suites -- unchanged -- run with
 * {@code CudaArrayKernel}, {@code CudaFftService} and {@code CudaImageKernel} bound behind the modal delegates.
 * <p>
 * It follows the shape of the reference's native drivers ({@code src_test/org/shared/test/AllNative.java:78-105} for
 * libsst, {@code src_test/org/sharedx/test/AllX.java:76-99} for the FFTW extension): load the library through the registry
 * class loader, switch the delegates to the registered services, run the suites. The only difference between a run of
 * {@code AllNative} / {@code AllX} and a run of this class is the library that is loaded -- that is the whole drop-in claim.
 * </p>
 * <p>
 * Not compiled in this repository's build image (no JDK); INTEGRATION.md lists the ant / javac line that builds it
 * against the reference's class path.
 * </p>
 */
package org.sharedx.cuda.test;

import org.shared.array.ArrayBase;
import org.shared.image.kernel.ImageOps;
import org.shared.log.Logging;
import org.shared.metaclass.Loader;
import org.shared.metaclass.Loader.EntryPoint;
import org.shared.metaclass.Loader.LoadableResources;
import org.shared.metaclass.RegistryClassLoader;
import org.shared.test.Tests;
import org.shared.test.array.AllArrayOperationTests;
import org.shared.test.fft.AllFftTests;
import org.shared.test.image.AllImageTests;
import org.shared.util.Control;

@LoadableResources(resources = { "jar:lib.junit", "jar:lib.log4j", "jar:lib.slf4j-api", "jar:lib.slf4j-log4j12" }, //
packages = { "org.sharedx.cuda.test", "org.shared.test", "org.shared.test.array", "org.shared.test.fft",
        "org.shared.test.image" })
public class AllCuda {

    /** Delegates to {@link Loader#run(String, Object)}, as {@code AllX.main} does. */
    public static void main(String[] args) throws Exception {
        Loader.run("org.sharedx.cuda.test.AllCuda", null);
    }

    /** The program entry point. */
    @EntryPoint
    public static void main0(String[] args) {

        Logging.configureLog4J("org/shared/log4j.xml");
        Logging.configureLog4J("org/shared/test/log4j.xml");

        try {

            // JNI_OnLoad registers the three providers (shared_b200/csrc/jni/jni_glue.cpp) -- or refuses to, without an
            // sm_100 device: there is no CPU fallback to fall through to.
            ((RegistryClassLoader) Thread.currentThread().getContextClassLoader()).loadLibrary("lib/sst_cuda");

        } catch (RuntimeException e) {

            Tests.log.debug("libsst_cuda not found or no CUDA device. Skipping tests.");

            return;
        }

        Control.checkTrue(ArrayBase.opKernel.useRegisteredKernel() //
                && ArrayBase.fftService.useRegisteredService() //
                && ImageOps.imKernel.useRegisteredKernel(), //
                "Could not link the CUDA providers");

        Tests.runTests("CUDA Provider Tests", //
                AllArrayOperationTests.class, // ArrayKernelTest, RealArrayTest, IntegerArrayTest, MatrixTest, SparseArrayTest ..
                AllFftTests.class, // ArrayFftTest, ConvolutionCacheTest
                AllImageTests.class, // IntegralImageTest, IntegralHistogramTest
                BenchmarkCuda.class);
    }

    // Dummy constructor.
    AllCuda() {
    }
}
