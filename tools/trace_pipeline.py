"""Dump the per-k-iteration SM-clock stamps of CTA 0 of one tcgen05 convolution (SGB_DEBUG_SKIP=16 [+ other bits])."""
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
t0 = t[0][0]
P = [("wait", 0), ("expect", 4), ("tmaA", 5), ("tmaB", 1)]
M = [("top", 7), ("wait", 8), ("fence", 9), ("elect", 2), ("mma", 10), ("commit", 3), ("syncw", 11)]
print("producer: stamp - previous stamp (cycles);  first column = cycles since the previous iteration's last stamp")
for i in range(1, 40):
    prev = t[1][i - 1]
    out = []
    for name, j in P:
        out.append(f"{name}+{t[j][i] - prev:4d}")
        prev = t[j][i]
    print(f"P it{i:3d} @ {t[0][i] - t0:7d}  " + "  ".join(out))
print()
for i in range(1, 40):
    prev = t[11][i - 1]
    out = []
    for name, j in M:
        out.append(f"{name}+{t[j][i] - prev:4d}")
        prev = t[j][i]
    print(f"M it{i:3d} @ {t[8][i] - t0:7d}  " + "  ".join(out))
