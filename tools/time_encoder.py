"""Encoder timing (CUDA events, L2 flushed between iterations): python tools/time_encoder.py [--batches 1,4] [--size 512]
Selects the implementation with CHORE_B200_ENCODER=hx|tc1|simt (read at import)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import chore_b200
from oracle import chore_oracle as O

ap = argparse.ArgumentParser()
ap.add_argument("--batches", default="1,4")
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--iters", type=int, default=20)
a = ap.parse_args()
dev = "cuda:0"
net = chore_b200.CHORE(device=dev)
net.load_state_dict(O.make_state_dict(0, "unit"))
h = net.handle
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = {"mode": os.environ.get("CHORE_B200_ENCODER", "hx"), "size": a.size}
for B in [int(x) for x in a.batches.split(",")]:
    img = O.synth_images(1, B=B, size=a.size).to(dev)
    for _ in range(3):
        h.encode(img)
    torch.cuda.synchronize()
    ts = []
    for _ in range(a.iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        keep = h.encode(img)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    out[f"B{B}_ms_median"] = ts[len(ts) // 2]
    out[f"B{B}_ms_min"] = ts[0]
    out[f"B{B}_ms_per_image"] = ts[len(ts) // 2] / B
    out[f"B{B}_tflops_algorithmic"] = 258.25e9 * B * (a.size / 512.0) ** 2 / (ts[len(ts) // 2] * 1e-3) / 1e12
print(json.dumps(out))
