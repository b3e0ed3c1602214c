#!/usr/bin/env python
"""Hand-off timeline of the layer-1 kernel (pair 0, leader CTA): clock64() stamps per tile, printed relative to the
first stamp in microseconds AT THE NOMINAL CLOCK nvidia-smi reports after the run (under the power cap the real SM clock
inside the kernel is lower: compare with the per-CTA globaltimer stamps printed below).
usage: probe_trace.py [rows=75776] [layer-1 variant (tuning 24)=2] [-] [debug flags: 8 = no TMA stores] [tuning 25 value]"""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402

from kubetorch_b200.device import lib as L  # noqa: E402
from kubetorch_b200.device import mlp, ops  # noqa: E402

ops.ensure_init([0])
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 75776
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 2
g = torch.Generator(device="cuda").manual_seed(0)
w1 = (torch.randn(1024, 256, device="cuda", generator=g) * 0.02).bfloat16()
w2 = (torch.randn(1024, 1024, device="cuda", generator=g) * 0.02).bfloat16()
w3 = (torch.randn(64, 1024, device="cuda", generator=g) * 0.02).bfloat16()
obs = torch.randn(rows, 256, device="cuda", generator=g).bfloat16()
ops.set_tuning(24, variant)
if len(sys.argv) > 5:
    ops.set_tuning(25, int(sys.argv[5]))
if len(sys.argv) > 4:
    L.call("ktb_debug_set_ptr", 1, int(sys.argv[4]))
y = mlp.mlp_forward(obs, w1, w2, w3)
for _ in range(20):
    mlp.mlp_forward(obs, w1, w2, w3, out=y)
torch.cuda.synchronize()
trace = torch.zeros(2048 + 8 * 160, dtype=torch.int64, device="cuda")
L.call("ktb_debug_set_ptr", 0, trace.data_ptr())
mlp.mlp_forward(obs, w1, w2, w3, out=y)
torch.cuda.synchronize()
L.call("ktb_debug_set_ptr", 0, None)
mhz = float(subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm", "--format=csv,noheader,nounits", "-i", "0"],
                           capture_output=True, text=True).stdout.strip() or 1900)
tg = trace.cpu()[1024:1024 + 4 * 148].view(148, 4).double()
t = trace.cpu()[:1024].view(16, 64)
t0 = int(t[t > 0].min())
names = {0: "mma_buf_free", 1: "mma_kb0_data", 12: "mma_kb1_data", 13: "mma_kb2_data", 2: "mma_kb3_data", 3: "mma_commit", 4: "epi_start", 10: "epi_h0_smem_free",
         5: "epi_h0_done", 11: "epi_h1_smem_free", 6: "epi_arrive", 7: "epi_end", 8: "prod_tile_begin", 9: "prod_tile_issued"}
order = [8, 9, 0, 1, 12, 13, 2, 3, 4, 10, 5, 11, 6, 7]
print("sm_mhz", mhz, "rows", rows, "variant", variant)
print("tile " + " ".join(f"{names[e]:>13s}" for e in order))
ntiles = int((t[4] > 0).sum())
for i in range(ntiles):
    print(f"{i:4d} " + " ".join(f"{(int(t[e, i]) - t0) / mhz:13.3f}" if int(t[e, i]) else f"{'-':>13s}" for e in order))
per = [(int(t[7, i]) - int(t[7, i - 1])) / mhz for i in range(2, ntiles)]
print(json.dumps({"tiles": ntiles, "period_us_mean": sum(per) / max(1, len(per)),
                  "epi_busy_us_mean": sum((int(t[7, i]) - int(t[4, i])) / mhz for i in range(ntiles)) / max(1, ntiles),
                  "mma_us_mean": sum((int(t[3, i]) - int(t[0, i])) / mhz for i in range(ntiles)) / max(1, ntiles),
                  "commit_to_epi_us_mean": sum((int(t[4, i]) - int(t[3, i])) / mhz for i in range(ntiles)) / max(1, ntiles)}))

g0 = float(tg[:, 0].min())
rel = (tg - g0) / 1000.0
print("per-CTA globaltimer (us after the first CTA start): start / setup done / work done / exit")
for k, nm in enumerate(("start", "setup_done", "work_done", "exit")):
    col = rel[:, k]
    print(f"  {nm:11s} min {float(col.min()):7.2f}  median {float(col.median()):7.2f}  max {float(col.max()):7.2f}")
print("  CTA 0:", [round(float(v), 2) for v in rel[0]], " CTA 147:", [round(float(v), 2) for v in rel[147]])
dur = rel[:, 2] - rel[:, 1]
print(f"  work duration per CTA: min {float(dur.min()):.2f} median {float(dur.median()):.2f} max {float(dur.max()):.2f}")

tw = (trace.cpu()[2048:2048 + 8 * 148].view(148, 8).double() - g0) / 1000.0
print("per-warp arrival at the final cluster barrier (us): warp 0 producer, 1 MMA, 2-5 epilogue")
for w in range(6):
    col = tw[:, w]
    print(f"  warp {w}: min {float(col.min()):7.2f} median {float(col.median()):7.2f} max {float(col.max()):7.2f}")
