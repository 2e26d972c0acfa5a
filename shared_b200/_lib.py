"""ctypes binding of libsst_cuda.so (the C ABI declared in include/sst_cuda.h).

The library is built in-tree by `__graft_entry__.build()` / `make -C shared_b200/csrc`. There is no CPU
fallback: if the shared object is missing, or no sm_100 device is present when a compute call is made,
the call raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsst_cuda.so")

_lib = None

c_i32_p = ctypes.POINTER(ctypes.c_int32)
c_i64_p = ctypes.POINTER(ctypes.c_int64)
c_dbl_p = ctypes.POINTER(ctypes.c_double)
c_void_pp = ctypes.POINTER(ctypes.c_void_p)

# name -> (restype, argtypes); one entry per symbol declared in include/sst_cuda.h
SIGNATURES = {
    "sstc_init": (ctypes.c_int, [ctypes.c_int]),
    "sstc_shutdown": (ctypes.c_int, []),
    "sstc_last_error": (ctypes.c_char_p, []),
    "sstc_abi_version": (ctypes.c_int, []),
    "sstc_launch_count": (ctypes.c_int64, []),
    "sstc_malloc": (ctypes.c_int, [c_void_pp, ctypes.c_int64]),
    "sstc_free": (ctypes.c_int, [ctypes.c_void_p]),
    "sstc_host_alloc": (ctypes.c_int, [c_void_pp, ctypes.c_int64]),
    "sstc_host_free": (ctypes.c_int, [ctypes.c_void_p]),
    "sstc_h2d": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]),
    "sstc_d2h": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]),
    "sstc_d2d": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]),
    "sstc_memset": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p]),
    "sstc_stream_sync": (ctypes.c_int, [ctypes.c_void_p]),
    "sstc_fft_lengths": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, c_i32_p, c_i64_p, c_i64_p]),
    "sstc_fft": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, c_i32_p, ctypes.c_void_p, ctypes.c_int64,
                                ctypes.c_void_p, ctypes.c_int64]),
    "sstc_fft_set_devices": (ctypes.c_int, [ctypes.c_int, c_i32_p]),
    "sstc_exp_host_copy": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                          ctypes.c_int64]),
    "sstc_host_last_stats": (ctypes.c_int, [c_dbl_p, ctypes.c_int]),
    "sstc_host_option": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int64]),
    "sstc_fft_plan_create": (ctypes.c_int, [c_void_pp, ctypes.c_int, ctypes.c_int, c_i32_p]),
    "sstc_fft_plan_execute": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "sstc_fft_plan_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "sstc_fft_plan_workspace": (ctypes.c_int64, [ctypes.c_void_p]),
    "sstc_fft_plan_launches": (ctypes.c_int, [ctypes.c_void_p]),
    "sstc_fft_plan_set_hint": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int64]),
    "sstc_fft_convolve_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                             c_i32_p, ctypes.c_void_p]),
    "sstc_fft_axis_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32,
                                         ctypes.c_int64, ctypes.c_int, ctypes.c_double, ctypes.c_void_p]),
    "sstc_fft_axis_strided_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32,
                                                 ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                                 ctypes.c_int64, ctypes.c_int, ctypes.c_double, ctypes.c_void_p]),
    "sstc_fft_axis_blocked_dev": (ctypes.c_int, [c_void_pp, ctypes.c_int, ctypes.c_int64, c_void_pp, ctypes.c_int,
                                                 ctypes.c_int64, ctypes.c_int64, ctypes.c_int32, ctypes.c_int64,
                                                 ctypes.c_int, ctypes.c_double, ctypes.c_void_p]),
    "sstc_fft_lines_scatter_dev": (ctypes.c_int, [ctypes.c_void_p, c_void_pp, ctypes.c_int, ctypes.c_int64,
                                                  ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                                  ctypes.c_int, ctypes.c_double, ctypes.c_int64, ctypes.c_void_p]),
    "sstc_fft_lines_scatter_ex_dev": (ctypes.c_int, [ctypes.c_void_p, c_void_pp, ctypes.c_int, ctypes.c_int64,
                                                     ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                                     ctypes.c_int, ctypes.c_double, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p]),
    "sstc_fft_planes_scatter_dev": (ctypes.c_int, [ctypes.c_void_p, c_void_pp, ctypes.c_int, ctypes.c_int32, ctypes.c_int32,
                                                   ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                                   ctypes.c_int, ctypes.c_double, ctypes.c_int64, ctypes.c_void_p]),
    "sstc_slab_create": (ctypes.c_int, [c_void_pp, c_i32_p, ctypes.c_int, ctypes.c_int, c_void_pp, c_void_pp, c_void_pp]),
    "sstc_slab_forward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, c_void_pp, ctypes.c_void_p]),
    "sstc_slab_forward_begin": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "sstc_slab_inverse": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, c_void_pp, ctypes.c_void_p]),
    "sstc_slab_set_option": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int64]),
    "sstc_slab_trace": (ctypes.c_int64, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int64]),
    "sstc_slab_graph_count": (ctypes.c_int64, [ctypes.c_void_p]),
    "sstc_slab_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "sstc_fft3d_sharded_create": (ctypes.c_int, [c_void_pp, c_i32_p, ctypes.c_int, c_i32_p]),
    "sstc_fft3d_sharded_forward": (ctypes.c_int, [ctypes.c_void_p, c_void_pp, c_void_pp, c_void_pp]),
    "sstc_fft3d_sharded_inverse": (ctypes.c_int, [ctypes.c_void_p, c_void_pp, c_void_pp, c_void_pp]),
    "sstc_fft3d_sharded_sync": (ctypes.c_int, [ctypes.c_void_p]),
    "sstc_fft3d_sharded_set_option": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int64]),
    "sstc_fft3d_sharded_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "sstc_fft3d_sharded_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "sstc_peer_barrier_dev": (ctypes.c_int, [c_void_pp, ctypes.c_int, ctypes.c_int, ctypes.c_uint64, ctypes.c_void_p]),
    "sstc_exp_peer_copy": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_int, ctypes.c_void_p]),
    "sstc_exp_set": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int64]),
    "sstc_ipc_export": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p]),
    "sstc_ipc_open": (ctypes.c_int, [ctypes.c_char_p, c_void_pp]),
    "sstc_ipc_close": (ctypes.c_int, [ctypes.c_void_p]),
}

_vp, _i, _i64, _d = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_double
_ARRAY_OPS = {
    "sstc_ruop": [_i, _d, _vp, _i64],
    "sstc_cuop": [_i, _d, _d, _vp, _i64],
    "sstc_iuop": [_i, ctypes.c_int32, _vp, _i64],
    "sstc_eop": [_i, _i, _vp, _vp, _vp, _i64],
    "sstc_convert": [_i, _i, _vp, _i64, _vp, _i64],
    "sstc_raop": [_i, _vp, _i64, _vp],
    "sstc_caop": [_i, _vp, _i64, _vp],
    "sstc_rrop": [_i, _vp, _i64, c_i32_p, c_i32_p, _vp, _i64, c_i32_p, c_i32_p, _i, c_i32_p, _i],
    "sstc_riop": [_i, _vp, _i64, c_i32_p, c_i32_p, _i, _vp, _i64, _i],
    "sstc_rdop": [_i, _vp, _i64, c_i32_p, c_i32_p, _i, _vp, _i64, c_i32_p, _i],
    "sstc_map": [_i, c_i32_p, _i, _vp, _i64, c_i32_p, c_i32_p, _vp, _i64, c_i32_p, c_i32_p, _i],
    "sstc_slice": [_i, c_i32_p, _i, _vp, _i64, c_i32_p, c_i32_p, _vp, _i64, c_i32_p, c_i32_p, _i],
    "sstc_mul": [_vp, _i64, _vp, _i64, _i, _i, _vp, _i64, _i],
    "sstc_diag": [_vp, _i64, _vp, _i64, _i, _i],
}
for _name, _args in _ARRAY_OPS.items():
    SIGNATURES[_name] = (ctypes.c_int, _args)
    SIGNATURES[_name + "_dev"] = (ctypes.c_int, _args + [ctypes.c_void_p])  # device twin: same arguments + stream
SIGNATURES["sstc_shift_dev"] = (ctypes.c_int, [_i, _vp, _vp, c_i32_p, _i, _i, c_i32_p, _vp])
SIGNATURES["sstc_fftshift_dev"] = (ctypes.c_int, [_vp, _vp, c_i32_p, _i, _i, _vp])
SIGNATURES["sstc_transpose_dev"] = (ctypes.c_int, [_i, _vp, _vp, c_i32_p, _i, _i, c_i32_p, _vp])
SIGNATURES["sstc_pad_dev"] = (ctypes.c_int, [_vp, _vp, c_i32_p, _i, c_i32_p, _vp])
SIGNATURES["sstc_srand"] = (ctypes.c_int, [ctypes.c_uint64])
SIGNATURES["sstc_invert"] = (ctypes.c_int, [_vp, _i64, _vp, _i64, _i])
SIGNATURES["sstc_invert_dev"] = (ctypes.c_int, [_vp, _i64, _vp, _i64, _i, _vp])
SIGNATURES["sstc_svd"] = (ctypes.c_int, [_vp, _i64, _i, _i, _vp, _i64, _vp, _i64, _vp, _i64, _i, _i])
SIGNATURES["sstc_svd_dev"] = (ctypes.c_int, [_vp, _i64, _i, _i, _vp, _i64, _vp, _i64, _vp, _i64, _i, _i, _vp])
SIGNATURES["sstc_eigs"] = (ctypes.c_int, [_vp, _i64, _vp, _i64, _vp, _i64, _i])
SIGNATURES["sstc_eigs_dev"] = (ctypes.c_int, [_vp, _i64, _vp, _i64, _vp, _i64, _i, _vp])
SIGNATURES["sstc_insert_sparse"] = (ctypes.c_int, [_i, _vp, _i64, c_i32_p, c_i32_p, c_i32_p, _i, _vp, _vp, _i64, _vp, _vp])
SIGNATURES["sstc_slice_sparse"] = (ctypes.c_int, [_i, c_i32_p, _i, _vp, _i64, c_i32_p, c_i32_p, c_i32_p, _vp, c_i32_p, _vp,
                                                  _vp, _i64, c_i32_p, c_i32_p, c_i32_p, _vp, c_i32_p, _vp, _i, _vp])
SIGNATURES["sstc_find"] = (ctypes.c_int, [_vp, _i64, c_i32_p, c_i32_p, _i, c_i32_p, _vp, _i64, c_i32_p])
SIGNATURES["sstc_find_dev"] = (ctypes.c_int, [_vp, _i64, c_i32_p, c_i32_p, _i, c_i32_p, _vp, c_i32_p, _vp])
_IMAGE_OPS = {
    "sstc_integral_image": [_vp, _i64, c_i32_p, c_i32_p, _vp, _i64, c_i32_p, c_i32_p, _i],
    "sstc_integral_histogram": [_vp, _i64, c_i32_p, c_i32_p, _i, _vp, _i64, _vp, _i64, c_i32_p, c_i32_p],
}
for _name, _args in _IMAGE_OPS.items():
    SIGNATURES[_name] = (ctypes.c_int, _args)
    SIGNATURES[_name + "_dev"] = (ctypes.c_int, _args + [ctypes.c_void_p])


class SstCudaError(RuntimeError):
    """What the JNI glue raises as java.lang.RuntimeException (native/src/shared/Common.cpp:173-178)."""


def lib():
    """Load libsst_cuda.so (once). Raises if it has not been built: there is no fallback implementation."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SstCudaError(
                "libsst_cuda.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C shared_b200/csrc`); shared_b200 has no CPU fallback")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(status):
    if status != 0:
        raise SstCudaError(lib().sstc_last_error().decode("utf-8", "replace"))


def i32_array(values):
    values = [int(v) for v in values]
    return (ctypes.c_int32 * len(values))(*values)


_initialised = set()


def init(device=0):
    """sstc_init once per device per process."""
    if device not in _initialised:
        check(lib().sstc_init(int(device)))
        _initialised.add(device)
