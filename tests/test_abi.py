"""CPU-side checks of the drop-in boundary: libsst_cuda.so loads, exports every symbol include/sst_cuda.h
declares, and refuses to compute without a CUDA device (no CPU fallback)."""
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "sst_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sstc_\w+)\s*\(", text)))


def test_header_declares_and_library_exports():
    from shared_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 20
    handle = _lib.lib()
    for n in names:
        assert hasattr(handle, n), "libsst_cuda.so does not export " + n
        assert n in _lib.SIGNATURES, "no ctypes signature for " + n
    assert handle.sstc_abi_version() == 2


def test_lengths_follow_plan_rules():
    # native/src/sharedx/Plan.cpp:230-320
    import shared_b200 as s
    assert s.transform_lengths(s.R2C, [4, 6]) == (24, 32)
    assert s.transform_lengths(s.C2R, [4, 7]) == (32, 28)
    assert s.transform_lengths(s.FORWARD, [3, 3, 3]) == (54, 54)
    assert s.transform_lengths(s.BACKWARD, [1024, 1024, 1024]) == (2 ** 31, 2 ** 31)  # beyond Java's int
    with pytest.raises(s.SstCudaError, match="Invalid dimensions"):
        s.transform_lengths(s.FORWARD, [4, 0])
    with pytest.raises(s.SstCudaError, match="Transform type not recognized"):
        s.transform_lengths(7, [4])


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import shared_b200 as s
    with pytest.raises(s.SstCudaError, match="no CPU fallback"):
        s.CudaFftService()


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "shared_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".java")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle|oracle/|libsst_oracle|libsst_ref", src, re.M), f


def test_jni_glue_covers_every_java_native():
    """Every `native` method of the Java providers has its JNI export in jni_glue.cpp, and every export has a native
    method (the glue cannot be linked here -- no JDK -- so this is what keeps the two sides from drifting apart)."""
    java_dir = os.path.join(ROOT, "shared_b200", "java", "org", "sharedx", "cuda")
    glue = open(os.path.join(ROOT, "shared_b200", "csrc", "jni", "jni_glue.cpp")).read()
    glue = glue.replace("AK(", "Java_org_sharedx_cuda_CudaArrayKernel_(").replace("CA(", "Java_org_sharedx_cuda_CudaComplexArray_(")
    exported = set(re.findall(r"Java_org_sharedx_cuda_(\w+?)_\(?(\w+)\)?\s*\(JNIEnv", glue))
    declared = set()
    for f in os.listdir(java_dir):
        src = re.sub(r"/\*.*?\*/", "", open(os.path.join(java_dir, f)).read(), flags=re.S)
        cls = f[:-5]
        for m in re.finditer(r"\bnative\b[^;(]*?(\w+)\s*\(", src):
            declared.add((cls, m.group(1)))
    assert len(declared) >= 20
    assert declared == exported, (sorted(declared - exported), sorted(exported - declared))


def test_python_mirrors_have_the_reference_method_names():
    """The host-side mirrors expose the SPI method names of the reference interfaces (ArrayKernel.java:285-674,
    FftService.java:50-105, ImageKernel.java:56-80) -- the parity tests read like the reference's own."""
    from shared_b200.array_kernel import CudaArrayKernel
    from shared_b200.fft_service import CudaFftService
    from shared_b200.image_kernel import CudaImageKernel
    for name in ("randomize derandomize map slice rrOp riOp rdOp raOp caOp ruOp cuOp iuOp eOp convert mul diag svd eigs "
                 "invert find insertSparse sliceSparse").split():
        assert callable(getattr(CudaArrayKernel, name)), name
    for name in "fft ifft rfft rifft setHint getHint".split():
        assert callable(getattr(CudaFftService, name)), name
    for name in "createIntegralImage createIntegralHistogram".split():
        assert callable(getattr(CudaImageKernel, name)), name


def test_tools_and_entry_points_compile():
    """bench.py, __graft_entry__.py and every driver under tools/ at least byte-compile (they only run on the GPU box)."""
    paths = [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]
    paths += [os.path.join(ROOT, "tools", f) for f in sorted(os.listdir(os.path.join(ROOT, "tools"))) if f.endswith(".py")]
    for p in paths:
        compile(open(p).read(), p, "exec")


def test_cmake_tree_matches_the_makefile():
    """native/CMakeLists.txt and shared_b200/csrc/Makefile build the same library: same sources (a glob over csrc/*.cu in
    both), and the same set of files compiled without FMA contraction -- bit-exact parity depends on that flag."""
    cm = open(os.path.join(ROOT, "native", "CMakeLists.txt")).read()
    mk = open(os.path.join(ROOT, "shared_b200", "csrc", "Makefile")).read()
    nofma_mk = set(re.search(r"^NOFMA := (.*)$", mk, re.M).group(1).split())
    block = re.search(r"set_source_files_properties\((.*?)PROPERTIES COMPILE_OPTIONS \"-fmad=false\"\)", cm, re.S).group(1)
    nofma_cm = set(re.findall(r"/(\w+)\.cu", block))
    assert nofma_mk == nofma_cm and len(nofma_mk) >= 4
    assert "100a" in cm and "exports.map" in cm and "jni_glue.cpp" in cm
    for name in nofma_mk:
        assert os.path.exists(os.path.join(ROOT, "shared_b200", "csrc", name + ".cu")), name


def test_jni_export_signatures_match_the_java_natives():
    """Type by type: every `native` method of the Java providers against the parameter list of its JNI export
    (double[] -> jdoubleArray, int -> jint, generic V / Object -> jobject, ...). A JVM resolves natives by name only, so
    a drifted parameter list would not fail at link time -- it would read garbage."""
    java_dir = os.path.join(ROOT, "shared_b200", "java", "org", "sharedx", "cuda")
    glue = open(os.path.join(ROOT, "shared_b200", "csrc", "jni", "jni_glue.cpp")).read()
    glue = re.sub(r"\bAK\((\w+)\)", r"Java_org_sharedx_cuda_CudaArrayKernel_\1", glue)
    glue = re.sub(r"\bCA\((\w+)\)", r"Java_org_sharedx_cuda_CudaComplexArray_\1", glue)
    jtype = {"double[]": "jdoubleArray", "int[]": "jintArray", "int": "jint", "double": "jdouble", "boolean": "jboolean",
             "long": "jlong", "void": "void", "V": "jobject", "int...": "jintArray", "Object": "jobject", "SparseArrayState<V>": "jobject"}
    exports = {}
    for m in re.finditer(r"JNIEXPORT\s+(\w+)\s+JNICALL\s+Java_org_sharedx_cuda_(\w+?)_(\w+)\(JNIEnv \*\w*,\s*jobject\w*(.*?)\)\s*\{",
                         glue, re.S):
        params = [re.sub(r"\s+\w+$", "", q.strip()) for q in m.group(4).split(",") if q.strip()]
        exports[(m.group(2), m.group(3))] = (m.group(1), params)
    checked = 0
    for f in sorted(os.listdir(java_dir)):
        src = re.sub(r"/\*.*?\*/|//[^\n]*", "", open(os.path.join(java_dir, f)).read(), flags=re.S)
        for m in re.finditer(r"\bnative\s+(?:<V>\s+)?([\w\[\]<>]+)\s+(\w+)\s*\((.*?)\)\s*;", src, re.S):
            ret, name, plist = m.group(1), m.group(2), m.group(3)
            params = [jtype[re.sub(r"\s+\w+$", "", q.strip())] for q in plist.split(",") if q.strip()]
            want = (jtype[ret], params)
            assert exports.get((f[:-5], name)) == want, (f, name, exports.get((f[:-5], name)), want)
            checked += 1
    assert checked >= 25


def test_sparse_argument_checks_need_no_device():
    """The reference validates before it touches array contents (SparseOpsInsert.cpp:77-108, SparseOpsSlice.cpp:87-166);
    so does the C ABI: these calls fail with the reference's texts before any CUDA call, hence also on a box without a GPU."""
    import ctypes

    import numpy as np

    from shared_b200 import _lib
    from shared_b200.array_kernel import SparseStateStruct
    lib = _lib.lib()

    def i32(v):
        return np.ascontiguousarray(v, dtype=np.int32)

    def pi(a):
        return a.ctypes.data_as(_lib.c_i32_p)

    def vp(a):
        return ctypes.c_void_p(a.ctypes.data)

    st = SparseStateStruct()
    vals, idx = np.array([1.0, 2.0]), i32([1, 3])
    empty_v, empty_i = np.zeros(1), i32([0])

    def insert(D, S, Do, elem_bytes=8):
        D, S, Do = i32(D), i32(S), i32(Do)
        return lib.sstc_insert_sparse(elem_bytes, vp(empty_v), 0, pi(D), pi(S), pi(Do), D.size, vp(empty_i), vp(vals), 2,
                                      vp(idx), ctypes.byref(st))

    def last():
        return lib.sstc_last_error().decode()

    assert insert([2, 3], [4, 1], [0, 3, 7]) == 1 and last() == "Invalid dimensions and/or strides"
    assert insert([2, -3], [3, 1], [0, 3, 7]) == 1 and last() == "Invalid dimensions and/or strides"
    assert insert([2, 3], [3, 1], [0, 3, 8]) == 1 and last() == "Invalid arguments"
    assert insert([2, 3], [3, 1], [0, 3, 7], elem_bytes=2) == 1 and last() == "Invalid array types"
    assert lib.sstc_insert_sparse(8, vp(empty_v), 0, pi(i32([2])), pi(i32([1])), pi(i32([0, 3])), 1, vp(empty_i), vp(vals), 2,
                                  vp(idx), None) == 1 and last() == "Invalid arguments"

    D, S, Do = i32([2, 3]), i32([3, 1]), i32([0, 3, 7])
    io, ii = i32(np.zeros(7)), i32([0])

    def slc(slices):
        sl = i32(slices)
        return lib.sstc_slice_sparse(8, pi(sl), sl.size, vp(empty_v), 0, pi(D), pi(S), pi(Do), vp(empty_i), pi(io), vp(ii),
                                     vp(empty_v), 0, pi(D), pi(S), pi(Do), vp(empty_i), pi(io), vp(ii), 2, ctypes.byref(st))

    assert slc([0, 0, 2]) == 1 and last() == "Invalid dimension"
    assert slc([2, 0, 0]) == 1 and last() == "Invalid index"
    assert slc([0, 3, 1]) == 1 and last() == "Invalid index"
    assert slc([0, 0]) == 1 and last() == "Invalid arguments"
