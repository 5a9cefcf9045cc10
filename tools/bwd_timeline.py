"""Protocol timeline of CTA 0 of dense_bwd_tc_kernel (SM clocks at the barrier points of each tile), Gowalla-sized
layer.  python tools/bwd_timeline.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from seoul_tourism_recommendation_ngcf_b200 import _lib

lib = _lib.load()
raw = C.CDLL(_lib.LIB_PATH)
raw.ngcf_debug_bwd_timeline.argtypes = [C.c_int, C.c_void_p]
dev = torch.device("cuda:0")
MESS_P = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
N, d = 70839, 64
st = torch.cuda.current_stream().cuda_stream
X = torch.randn(N, d, device=dev); S = torch.randn(N, d, device=dev)
E_out = torch.randn(N, d, device=dev); gE = torch.randn(N, d, device=dev)
W1 = torch.randn(d, d, device=dev) * 0.1; W2 = torch.randn(d, d, device=dev) * 0.1
gS = torch.empty(N, d, device=dev); gEl = torch.empty(N, d, device=dev); gM = torch.empty(N, d, device=dev)
gW1 = torch.zeros(d, d, device=dev); gW2 = torch.zeros(d, d, device=dev); gb1 = torch.zeros(d, device=dev); gb2 = torch.zeros(d, device=dev)
slot = torch.full((N,), -1, dtype=torch.int32, device=dev)
rows = torch.randperm(N, device=dev)[:3000]; slot[rows] = torch.arange(3000, dtype=torch.int32, device=dev)
gsum = torch.randn(3000, 4 * d, device=dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

def run():
    lib.ngcf_dense_bwd(gE.data_ptr(), slot.data_ptr(), gsum.data_ptr(), 4 * d, d, E_out.data_ptr(), S.data_ptr(), X.data_ptr(), N, d, d, W1.data_ptr(), W2.data_ptr(), 0.2, None, None, MESS_P, 1, None, 0, 0, 1, gS.data_ptr(), gEl.data_ptr(), gW1.data_ptr(), gb1.data_ptr(), gW2.data_ptr(), gb2.data_ptr(), gM.data_ptr(), st)

for _ in range(3): run()
torch.cuda.synchronize()
raw.ngcf_debug_bwd_timeline(1, None)
flush.zero_(); run(); torch.cuda.synchronize()
out = np.zeros(4 * 8 * 8, dtype=np.int64)
raw.ngcf_debug_bwd_timeline(0, out.ctypes.data)
t = out.reshape(4, 8, 8)
t0 = t[0, 7, 7]
print("all clocks relative to the end of the CTA's setup, in cycles; total =", t[1, 7, 7] - t0)
names = {0: ["start", "ES filled", "got tmem_full", "combined", "streamed out"],
         1: ["wait full_gm", "got full_gm", "got tmem_empty", "issued"],
         2: ["start", "loads issued", "got empty_gm", "gM done"]}
for role, rn in ((2, "loader"), (1, "mma"), (0, "epilogue")):
    for it in range(5):
        row = [(n, int(t[role, it, k] - t0)) for k, n in enumerate(names[role]) if t[role, it, k] > 0]
        if row:
            print(f"{rn:9s} tile {it}: " + "  ".join(f"{n}={v}" for n, v in row))
for it in range(4):
    print(f"loader tile {it}: half 0 computed / next half issued / half 1 computed / fenced:", [int(v - t0) for v in t[3, it]])

# weight-gradient kernel (enable = 2 stamps wgrad_tc_kernel instead)
raw.ngcf_debug_bwd_timeline(2, None)
flush.zero_(); run(); torch.cuda.synchronize()
out = np.zeros(4 * 8 * 8, dtype=np.int64)
raw.ngcf_debug_bwd_timeline(0, out.ctypes.data)
t = out.reshape(4, 8, 8); t0 = t[0, 7, 7]
print("wgrad_tc_kernel CTA 0, cycles after set-up: total =", int(t[1, 7, 7] - t0),
      " loader stage ends:", [int(t[2, it, 0] - t0) for it in range(8) if t[2, it, 0] > 0],
      " flush:", int(t[0, 0, 0] - t0), "->", int(t[0, 0, 1] - t0))
