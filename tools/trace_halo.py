"""Per-tile SM-clock timeline of CTA 0 of the halo-tile 3x3 convolution (SGB_DEBUG_SKIP=16)."""
import ctypes
import os
import sys

os.environ["SGB_DEBUG_SKIP"] = str(int(os.environ.get("SGB_DEBUG_SKIP", "0")) | 16)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from super_gradients_b200 import kernels as K  # noqa: E402
from super_gradients_b200 import lib  # noqa: E402

n, c, h, w, k, r, s = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "32,32,160,160,32,3,1").split(","))
x = torch.randn(n, c, h, w, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
wt = torch.randn(k, c, r, r, device="cuda") * 0.05
krsc, _ = K.weight_prepare(wt)
for _ in range(2):
    y = K.conv_fprop(x, krsc, k, r, r, s, r // 2)
torch.cuda.synchronize()
buf = (ctypes.c_int64 * 6144)()
lib.call("sgb_debug_read_trace", buf)
t = [list(buf[i * 512:(i + 1) * 512]) for i in range(12)]
t0 = t[7][0]
print("tile | P: top a_empty_ok A_issued | M: top acc_empty_ok a_full_ok mma_done | E: top acc_full_ok stored   (cycles since start)")
ntile = sum(1 for i in range(512) if 0 < t[7][i] - t0 < 10**9 or i == 0)
for i in range(0, min(ntile, 24)):
    v = lambda j: t[j][i] - t0
    print(f"{i:3d} | P {v(7):7d} {v(0):7d} {v(1):7d} | M {v(8):7d} {v(2):7d} {v(3):7d} {v(4):7d} | E {v(9):7d} {v(5):7d} {v(6):7d}")
