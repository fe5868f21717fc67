#!/usr/bin/env python
"""Query-point throughput sweep (BASELINE.json configs[4]): N points x B images on one GPU, forward and
forward+backward, with the achieved fp32-equivalent TFLOP/s of the fused query kernel.  Writes a CSV table
(`--out`, default profiles/sweep.csv).  Not part of the bench.py contract; a measurement aid."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "sweep.csv"))
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    import chore_b200
    from oracle import chore_oracle as O          # input / weight synthesis only
    dev = "cuda:0"
    net = chore_b200.CHORE(device=dev)
    net.load_state_dict(O.make_state_dict(0, "unit"))
    rows = ["B,N,fwd_ms,fwd_Mpts_per_s,fwd_TFLOPs_fp32_equiv,fwdbwd_df_ms,fwdbwd_Mpts_per_s"]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for B in (1, 8, 32):
        feat, tmpx = O.synth_features(B, B=B)
        feat, skip = feat.to(dev).permute(0, 2, 3, 1).contiguous(), tmpx.to(dev).permute(0, 2, 3, 1).contiguous()
        cc = torch.tensor([[1008., 995.]], device=dev).repeat(B, 1)
        for N in (1000, 20000, 100000, 1000000):
            if B * N > 8_000_000:
                continue
            pts = O.synth_points("init_box", N, B, N).to(dev)
            g_df = torch.ones(B, 2, N, device=dev)

            def timed(fn):
                for _ in range(3):
                    fn()
                ms = []
                for _ in range(args.reps):
                    flush.zero_()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(); fn(); b.record()
                    torch.cuda.synchronize()
                    ms.append(a.elapsed_time(b))
                return sorted(ms)[len(ms) // 2]

            f = timed(lambda: net.handle.query_fwd(feat, skip, pts, cc, 15))
            fb = timed(lambda: (net.handle.query_fwd(feat, skip, pts, cc, 1), net.handle.query_bwd(feat, skip, pts, cc, [g_df, None, None, None])))
            rows.append(f"{B},{N},{f:.4f},{B * N / f / 1e3:.2f},{B * N * 600832 / f / 1e9:.2f},{fb:.4f},{B * N / fb / 1e3:.2f}")
            print(rows[-1], flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as fo:
        fo.write("# query kernel sweep on 1x B200 (median of %d, L2 flushed between reps); init-box points (~24 %% in image)\n" % args.reps)
        fo.write("\n".join(rows) + "\n")


if __name__ == "__main__":
    main()
