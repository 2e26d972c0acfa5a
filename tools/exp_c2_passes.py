"""Per-pass timing of BASELINE configs[1] (4096 x 4096 rfft / rifft) on one B200: CUDA events around each pass."""
import json
import sys

import torch

sys.path.insert(0, ".")
import shared_b200 as s

st = torch.cuda.current_stream().cuda_stream
s.init(0)
n, h = 4096, 2049


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
    return round(a.elapsed_time(b) / reps, 4)


r = torch.rand(n * n, dtype=torch.float64, device="cuda") - 0.5
spec = torch.empty(n * h, 2, dtype=torch.float64, device="cuda")
spec2 = torch.empty_like(spec)
back = torch.empty(n * n, dtype=torch.float64, device="cuda")
res = {}
# r2c rows only: a plan over [n] batched is not exposed; use dims [1, n]-style batches through the 2-D plan with d0 = 1
rows = s.FftPlan(s.R2C, [n * n // n, n]) if False else None
p1 = s.FftPlan(s.R2C, [n, n])
res["rfft_total"] = ev(lambda: p1.execute(r.data_ptr(), spec.data_ptr(), st))
res["c2c_columns_4096_inner2049_inplace"] = ev(lambda: s.fft_axis(spec.data_ptr(), spec.data_ptr(), 1, n, h, False, 1.0, st))
res["c2c_columns_4096_inner2049_outofplace"] = ev(lambda: s.fft_axis(spec.data_ptr(), spec2.data_ptr(), 1, n, h, False, 1.0, st))
res["c2c_columns_4096_inner2048"] = ev(lambda: s.fft_axis(spec.data_ptr(), spec.data_ptr(), 1, n, 2048, False, 1.0, st))
res["c2c_rows_4096_contig_x2049"] = ev(lambda: s.fft_axis(spec.data_ptr(), spec.data_ptr(), h, n, 1, False, 1.0, st))
res["c2c_columns_2048_inner4096"] = ev(lambda: s.fft_axis(spec.data_ptr(), spec.data_ptr(), 1, 2048, 4096, False, 1.0, st))
res["c2c_columns_1024_inner8192"] = ev(lambda: s.fft_axis(spec.data_ptr(), spec.data_ptr(), 1, 1024, 8192, False, 1.0, st))
# candidates for the long strided axis: a padded, 128-byte-aligned row pitch on the load side, and the column
# axis as 4 x 1024 (the radix-4 step folded into the r2c pass leaves four 1024-point column transforms, stored interleaved)
pad = 2056
wsp = torch.empty(n * pad, 2, dtype=torch.float64, device="cuda")
res["c2c_columns_4096_padded_pitch_2056_to_dense"] = ev(lambda: s.fft_axis_strided(
    wsp.data_ptr(), spec.data_ptr(), 1, n, h, n * pad, pad, n * h, h, False, 1.0, st))
res["c2c_columns_4096_padded_pitch_2056_inplace"] = ev(lambda: s.fft_axis_strided(
    wsp.data_ptr(), wsp.data_ptr(), 1, n, h, n * pad, pad, n * pad, pad, False, 1.0, st))
res["c2c_columns_4x1024_padded_in_interleaved_dense_out"] = ev(lambda: s.fft_axis_strided(
    wsp.data_ptr(), spec.data_ptr(), 4, 1024, h, 1024 * pad, pad, h, 4 * h, False, 1.0, st))
res["c2c_columns_4x1024_dense_in_interleaved_dense_out"] = ev(lambda: s.fft_axis_strided(
    spec2.data_ptr(), spec.data_ptr(), 4, 1024, h, 1024 * h, h, h, 4 * h, False, 1.0, st))
res["c2c_columns_8x512_padded_in_interleaved_dense_out"] = ev(lambda: s.fft_axis_strided(
    wsp.data_ptr(), spec.data_ptr(), 8, 512, h, 512 * pad, pad, h, 8 * h, False, 1.0, st))
res["c2c_columns_2x2048_padded_in_interleaved_dense_out"] = ev(lambda: s.fft_axis_strided(
    wsp.data_ptr(), spec.data_ptr(), 2, 2048, h, 2048 * pad, pad, h, 2 * h, False, 1.0, st))
# persistent grid for the tile that fills an SM (one 1024-thread CTA): takes the CTA turnover out of every tile
from shared_b200 import _lib
for ctas in (148, 296, 444):
    _lib.check(_lib.lib().sstc_exp_set(b"long_strided_ctas", ctas))
    res["c2c_columns_4096_inner2049_inplace_persistent_%d" % ctas] = ev(lambda: s.fft_axis(spec.data_ptr(), spec.data_ptr(), 1, n, h, False, 1.0, st))
    res["c2c_columns_2048_inner4096_persistent_%d" % ctas] = ev(lambda: s.fft_axis(spec.data_ptr(), spec.data_ptr(), 1, 2048, 4096, False, 1.0, st))
    res["rfft_total_persistent_%d" % ctas] = ev(lambda: p1.execute(r.data_ptr(), spec.data_ptr(), st))
_lib.check(_lib.lib().sstc_exp_set(b"long_strided_ctas", 0))
p2 = s.FftPlan(s.C2R, [n, n])
res["rifft_total"] = ev(lambda: p2.execute(spec.data_ptr(), back.data_ptr(), st))
res["bytes_c2c_pass_MB"] = 2 * 16 * n * h / 1e6
res["roof_us_c2c_pass_at_6545"] = 2 * 16 * n * h / 6545e3
print(json.dumps(res))
