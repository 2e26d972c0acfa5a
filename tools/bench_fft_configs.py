"""BASELINE.json configs[0] and configs[1] on one B200 (device tier, CUDA events): 256x256 c2c round trip and
4096x4096 rfft / rifft / FFT-based valid-region correlation with a 31x31 kernel. Checks results against
torch.fft (cuFFT) as a same-box yardstick and prints one JSON object."""
import json
import math
import sys

import torch

sys.path.insert(0, ".")
import shared_b200 as s
from shared_b200.array_kernel import COMPLEX, DeviceArrayKernel

peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if __import__("os").path.exists("MEASURED_PEAKS.json") else 6650.0
st = torch.cuda.current_stream().cuda_stream
dk = DeviceArrayKernel(0, st)


def ev(fn, reps=20):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


res = {"peak_gbs": peak}
# ---- C1: 256 x 256 c2c fft + ifft ----------------------------------------------------------------------
n = 256 * 256
x = torch.rand(n, 2, dtype=torch.float64, device="cuda") - 0.5
y, z = torch.empty_like(x), torch.empty_like(x)
f, b = s.FftPlan(s.FORWARD, [256, 256]), s.FftPlan(s.BACKWARD, [256, 256])
ms = ev(lambda: (f.execute(x.data_ptr(), y.data_ptr(), st), b.execute(y.data_ptr(), z.data_ptr(), st)), 200)
err = (torch.linalg.norm(z - x) / torch.linalg.norm(x)).item()
res["C1_256x256_fft_ifft"] = {"ms": round(ms, 5), "gflops": round(2 * 5 * n * math.log2(n) / ms / 1e6, 1),
                              "roundtrip_rel_l2": err}

# ---- C2: 4096 x 4096 rfft, rifft, correlation --------------------------------------------------------------
d = [4096, 4096]
n = 4096 * 4096
h = 4096 * 2049
r = torch.rand(n, dtype=torch.float64, device="cuda") - 0.5
spec = torch.empty(h, 2, dtype=torch.float64, device="cuda")
back = torch.empty(n, dtype=torch.float64, device="cuda")
rf, ri = s.FftPlan(s.R2C, d), s.FftPlan(s.C2R, d)
ms_f = ev(lambda: rf.execute(r.data_ptr(), spec.data_ptr(), st))
ref = torch.view_as_real(torch.fft.rfft2(r.view(4096, 4096))).reshape(h, 2)
e_f = (torch.linalg.norm(spec - ref) / torch.linalg.norm(ref)).item()
ms_i = ev(lambda: ri.execute(spec.data_ptr(), back.data_ptr(), st))
e_i = (torch.linalg.norm(back - r) / torch.linalg.norm(r)).item()
bytes_model = 8 * n + 16 * h + 2 * 16 * h  # r2c pass (read reals, write half-spectrum) + one c2c pass in place
flops = 2.5 * n * math.log2(n)
res["C2_rfft_4096"] = {"ms": round(ms_f, 4), "gflops": round(flops / ms_f / 1e6, 1), "model_gbs": round(bytes_model / ms_f / 1e6, 1),
                       "frac_of_measured_hbm": round(bytes_model / ms_f / 1e6 / peak, 3), "rel_l2_vs_cufft": e_f}
bytes_model_i = 3 * 16 * h + 16 * h + 8 * n  # c2c pass in -> workspace... then c2r
res["C2_rifft_4096"] = {"ms": round(ms_i, 4), "gflops": round(flops / ms_i / 1e6, 1),
                        "model_gbs": round(bytes_model / ms_i / 1e6, 1), "roundtrip_rel_l2": e_i}
ms_c = ev(lambda: torch.fft.rfft2(r.view(4096, 4096)), 10)
res["C2_rfft_4096_cufft_yardstick_ms"] = round(ms_c, 4)
# correlation (ConvolutionCache.convolve, ConvolutionCache.java:99-138): kernel FFT cached, image spectrum given
ker = torch.zeros(4096, 4096, dtype=torch.float64, device="cuda")
ker[:31, :31] = torch.rand(31, 31, dtype=torch.float64, device="cuda")
kerf = torch.empty(h, 2, dtype=torch.float64, device="cuda")
rf.execute(ker.data_ptr(), kerf.data_ptr(), st)
kerf[:, 1] *= -1.0  # uConj
prod = torch.empty_like(spec)


def convolve():
    dk.eOp(2, COMPLEX, spec.data_ptr(), kerf.data_ptr(), prod.data_ptr(), 2 * h)  # CE_MUL
    ri.execute(prod.data_ptr(), back.data_ptr(), st)


ms_conv = ev(convolve)
want = torch.fft.irfft2(torch.view_as_complex(spec).view(4096, 2049) * torch.view_as_complex(kerf).view(4096, 2049),
                        s=(4096, 4096))
convolve()
e_c = (torch.linalg.norm(back.view(4096, 4096)[:4066, :4066] - want[:4066, :4066]) / torch.linalg.norm(want)).item()
import ctypes
import numpy as np
from shared_b200 import _lib
crop = torch.empty(4066 * 4066, dtype=torch.float64, device="cuda")
od = np.array([4066, 4066], dtype=np.int32)


def convolve_fused():
    _lib.check(_lib.lib().sstc_fft_convolve_dev(ri._h, ctypes.c_void_p(spec.data_ptr()), ctypes.c_void_p(kerf.data_ptr()),
                                                ctypes.c_void_p(crop.data_ptr()), od.ctypes.data_as(_lib.c_i32_p),
                                                ctypes.c_void_p(st)))


def convolve_3step():
    convolve()
    dk.map(8, [0, 0, 4066, 0, 0, 4066], back.data_ptr(), n, [4096, 4096], [4096, 1], crop.data_ptr(), 4066 * 4066,
           [4066, 4066], [4066, 1])


ms_3 = ev(convolve_3step)
ms_fused = ev(convolve_fused)
convolve_fused()
e_f2 = (torch.linalg.norm(crop.view(4066, 4066) - want[:4066, :4066]) / torch.linalg.norm(want[:4066, :4066])).item()
res["C2_correlation_31x31"] = {"ms_eMul_plus_rifft": round(ms_conv, 4), "ms_three_step_incl_crop": round(ms_3, 4),
                               "ms_fused_sstc_fft_convolve": round(ms_fused, 4), "rel_l2_vs_cufft": e_c,
                               "fused_rel_l2_vs_cufft": e_f2,
                               "note": "three-step = eMul kernel + rifft + subarray map (the reference's structure); "
                                       "fused = product on load of the first inverse pass, crop on store of the c2r pass"}
print(json.dumps(res))
