#!/usr/bin/env python
"""Bandwidth of the BatchNorm / QARepVGG elementwise passes at YOLO-NAS-S shapes (CUDA events, L2 flushed)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from super_gradients_b200 import kernels as K  # noqa: E402

SHAPES = [(32, 48, 320, 320), (32, 96, 160, 160), (32, 32, 160, 160), (32, 64, 80, 80), (32, 192, 80, 80), (32, 96, 40, 40)]


def bench(fn, flush, iters=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts) * 1e3


def main():
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for n, c, h, w in SHAPES:
        t = lambda: torch.randn(n, c, h, w, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)  # noqa: E731
        y3, u, dout = t(), t(), t()
        one, zero = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
        rm, rv = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
        tb = y3.numel() * 2 / 1e3  # KB per tensor -> GB/s = KB / us * 1e-3 ... use bytes / us / 1e3
        tbytes = y3.numel() * 2

        def gbs(us, ntens):
            return ntens * tbytes / us / 1e3

        out, coef = K.qarep_fwd(y3, u, one, zero, zero, one, zero, rm, rv, rm.clone(), rv.clone(), 1e-3, 1e-3, 0.03, "relu")
        d = K.qarep_desc(y3, u, out, 1e-3, 1e-3, 0.03, "relu", True)
        import ctypes
        from super_gradients_b200 import lib as L

        mom = torch.zeros((5, c), dtype=torch.float64, device="cuda")
        us = bench(lambda: L.call("sgb_qarep_moments", ctypes.byref(d), K._ptr(y3), K._ptr(u), K._ptr(mom), K._stream()), flush)
        print(f"C={c:3d} {h}x{w} qarep_moments    {us:8.1f} us {gbs(us, 2):7.0f} GB/s")
        us = bench(lambda: K.qarep_fwd(y3, u, one, zero, zero, one, zero, rm, rv, rm, rv, 1e-3, 1e-3, 0.03, "relu"), flush)
        print(f"C={c:3d} {h}x{w} moments+fwd      {us:8.1f} us {gbs(us, 5):7.0f} GB/s")
        us = bench(lambda: K.qarep_bwd(dout, out, y3, u, coef, one, one, 1e-3, 1e-3, "relu"), flush)
        print(f"C={c:3d} {h}x{w} qarep bwd (2 pass){us:8.1f} us {gbs(us, 8):7.0f} GB/s")
        stats = K.channel_stats(y3)
        us = bench(lambda: K.channel_stats(y3), flush)
        print(f"C={c:3d} {h}x{w} channel_stats    {us:8.1f} us {gbs(us, 1):7.0f} GB/s")
        y, mean, rstd = K.bn_act_fwd(y3, stats, one, zero, rm, rv, 1e-3, 0.03, "relu")
        us = bench(lambda: K.bn_act_fwd(y3, stats, one, zero, rm, rv, 1e-3, 0.03, "relu"), flush)
        print(f"C={c:3d} {h}x{w} bn_act_fwd       {us:8.1f} us {gbs(us, 2):7.0f} GB/s")
        us = bench(lambda: K.bn_act_bwd(dout, y3, y, one, mean, rstd, 1e-3, "relu", beta=zero), flush)
        print(f"C={c:3d} {h}x{w} bn bwd (2 pass)  {us:8.1f} us {gbs(us, 5):7.0f} GB/s")


if __name__ == "__main__":
    main()
