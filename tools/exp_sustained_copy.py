"""Burst against sustained HBM copy rate (the roofline denominator): b.copy_(a) over 2 GiB, timed in blocks of 50 copies for
about 3 s, with nvidia-smi clocks / power sampled alongside. One JSON object."""
import json
import subprocess
import threading
import time

import torch

a = torch.empty(1 << 28, dtype=torch.float64, device="cuda")  # 2 GiB
b = torch.empty_like(a)
a.normal_()
rows = []
proc = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap",
                         "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [rows.append((time.time(), l.strip())) for l in proc.stdout], daemon=True).start()
for _ in range(5):
    b.copy_(a)
torch.cuda.synchronize()
blocks = []
t0 = time.time()
while time.time() - t0 < 3.0:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        b.copy_(a)
    e1.record()
    torch.cuda.synchronize()
    blocks.append(round(2 * a.numel() * 8 * 50 / e0.elapsed_time(e1) / 1e6, 1))
t1 = time.time()
proc.terminate()
inside = [r for (t, r) in rows if t0 <= t <= t1]
print(json.dumps({"what": "b.copy_(a), 2 GiB float64, read + write bytes, GB/s per block of 50 copies over 3 s",
                  "gbs_first_block": blocks[0], "gbs_last_block": blocks[-1], "gbs_min": min(blocks), "gbs_max": max(blocks),
                  "blocks": len(blocks), "smi_first": inside[:2], "smi_last": inside[-2:]}))
