"""Per-phase clock64 stamps of CTA 0 of every conv_hx launch of one encode (debug aid).  CHORE_B200_ENCODER_GRAPH=0 required."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["CHORE_B200_ENCODER_GRAPH"] = "0"
import torch
import chore_b200
from chore_b200 import _lib
from oracle import chore_oracle as O
dev = "cuda:0"
net = chore_b200.CHORE(device=dev)
net.load_state_dict(O.make_state_dict(0, "unit"))
h = net.handle
img = O.synth_images(1, B=1, size=512).to(dev)
for _ in range(2):
    h.encode(img)
torch.cuda.synchronize()
buf = torch.zeros(256 * 32, dtype=torch.int64, device=dev)
lib = _lib.load_library()
lib.chore_debug_hx_trace.argtypes = [ctypes.c_void_p]
lib.chore_debug_hx_trace(buf.data_ptr())
h.encode(img)
torch.cuda.synchronize()
lib.chore_debug_hx_trace(None)
t = buf.view(256, 32).cpu()
names = {0: "start", 1: "setup", 12: "ld_issued", 13: "scsh", 2: "table", 14: "tx_go", 3: "tx0", 4: "tx1", 5: "tx2", 6: "tx3", 8: "mma0", 9: "mma1", 10: "mma2", 11: "mma3",
         16: "accfull", 19: "parts", 20: "c0_ld", 21: "c0_raw", 22: "c0_straw", 23: "c0_out", 25: "c0_stout", 17: "epi", 18: "flush", 24: "end"}
for i in range(256):
    r = t[i]
    if r[0] == 0:
        break
    print(i, " ".join(f"{names[k]}={int(r[k] - r[0])}" for k in names if r[k] != 0 and k != 0))
