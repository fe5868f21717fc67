import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
import chore_b200
from chore_b200 import _lib
from oracle import chore_oracle as O
dev = "cuda:0"
sd = O.make_state_dict(0, "unit")
net = chore_b200.CHORE(device=dev)
net.load_state_dict(sd)
feat, tmpx = O.synth_features(3, B=1)
net.im_feat_list, net.tmpx = [feat.to(dev)], tmpx.to(dev)
N = int(os.environ.get("NPTS", 4194304))
pts = O.synth_points("frustum", 2, 1, N).to(dev)
cc = torch.tensor([[1008., 995.]], device=dev)
f, s = net._maps()
h = net.handle
for mask in (15,):
    for _ in range(2):
        h.query_fwd(f, s, pts, cc, mask)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        h.query_fwd(f, s, pts, cc, mask)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"mask {mask}: {ms:.3f} ms  {N/ms/1e3:.1f} Mpts/s", flush=True)
