"""Test infrastructure: runs bench.py's `run_ours` control flow on CPU ranks (gloo) -- the CUDA runtime objects it touches are
replaced by inert fakes, the kernel wrappers by tests/cpu_backend.py, the CUDA graph by a replay of the eager step, and the
workload by a 32x32 image batch.  It exists to catch what cost round 1 its GPU budget: a collective that only some ranks
reach (a hang at N > 1), or an exception on a path that only runs at N > 1.  Launched by tests/test_bench_dryrun.py under
torchrun; usage: bench_dryrun.py REPO_ROOT [bench.py flags]."""
import contextlib
import os
import sys
import time

root = sys.argv[1]
sys.path[:0] = [root, os.path.join(root, "tests")]
sys.argv = ["bench.py"] + sys.argv[2:]

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from _pytest.monkeypatch import MonkeyPatch  # noqa: E402

import bench  # noqa: E402
import cpu_backend  # noqa: E402
from super_gradients_b200 import functional as SF  # noqa: E402
from super_gradients_b200 import kernels as K  # noqa: E402
from super_gradients_b200 import lib as L  # noqa: E402
from super_gradients_b200.training import sg_trainer  # noqa: E402

mp = MonkeyPatch()
cpu_backend.install_training(mp)


class FakeEvent:
    def __init__(self, enable_timing=False):
        self.t = time.perf_counter()

    def record(self, stream=None):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return max((other.t - self.t) * 1e3, 1e-3)

    def synchronize(self):
        pass


class FakeStream:
    cuda_stream = 0

    def wait_stream(self, s):
        pass

    def wait_event(self, e):
        pass

    def synchronize(self):
        pass


def setup_device(device=None):
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and not dist.is_initialized():
        dist.init_process_group("gloo", init_method="env://")
    return torch.device("cpu")


def fake_capture_region(self, fn, pool=None):
    """Stand-in for TrainStep._capture_region: like a real capture it executes nothing now; replay() runs fn() and copies its
    outputs into the static output tensors handed back here.  TrainStep.capture's own logic (warm-up, one graph or the
    data-parallel split around the all-reduce) runs unmodified on top of it."""
    template = (torch.zeros(()), torch.zeros(4))

    class FakeGraph:
        def replay(self):
            out = fn()
            if out is not None:
                template[0].copy_(out[0])
                template[1].copy_(out[1])

        def pool(self):
            return None

    return FakeGraph(), template


def synth_batch(batch, seed, img=32):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, 3, img, img, generator=g)
    rows = []
    for b in range(batch):
        cxy = torch.rand(bench.NBOX, 2, generator=g) * img * 0.5 + img * 0.25
        wh = torch.rand(bench.NBOX, 2, generator=g) * img * 0.4 + 6
        cls = torch.randint(0, bench.NCLS, (bench.NBOX, 1), generator=g).float()
        rows.append(torch.cat([torch.full((bench.NBOX, 1), float(b)), cls, cxy, wh], 1))
    return x, torch.cat(rows)


def profiled(name, fn):
    """The stand-ins bypass kernels._timed: record a profile entry for the conv family like the real wrappers do."""

    def wrapper(*a, **k):
        if not K.PROFILE_ON[0]:
            return fn(*a, **k)
        e0 = FakeEvent()
        out = fn(*a, **k)
        K.PROFILE.append((name, e0, FakeEvent(), ()))
        return out

    return wrapper


refuse = L.call
mp.setattr(L, "call", lambda name, *a: 0 if name == "sgb_check_device" else refuse(name, *a))
for n in ("conv_fprop", "conv_dgrad", "conv_wgrad"):
    mp.setattr(K, n, profiled("sgb_" + n, getattr(K, n)))
mp.setattr(sg_trainer, "setup_device", setup_device)
mp.setattr(sg_trainer.TrainStep, "_capture_region", fake_capture_region)
mp.setattr(bench, "synth_batch", synth_batch)
mp.setitem(bench.CONFIGS[2], "img", 32)  # the stand-in runs the graph of a 32 x 32 input
mp.setattr(torch.cuda, "Event", FakeEvent)
mp.setattr(torch.cuda, "Stream", FakeStream)
mp.setattr(torch.cuda, "current_stream", lambda *a: FakeStream())
mp.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
mp.setattr(torch.cuda, "synchronize", lambda *a: None)
mp.setattr(torch.cuda.profiler, "start", lambda: None)
mp.setattr(torch.cuda.profiler, "stop", lambda: None)

if os.environ.get("SGB_DRYRUN_THROTTLE") == "1":

    class ThrottledOnce:
        """Reports a hardware slowdown for the first timed region on rank 0 only (the other ranks see clean clocks)."""

        calls = 0

        def __init__(self, index, enabled=True):
            pass

        def start(self):
            pass

        def mark(self):
            pass

        def stop(self):
            pass

        def snapshot(self):
            ThrottledOnce.calls += 1
            hot = ThrottledOnce.calls == 1 and os.environ.get("RANK", "0") == "0"
            return {"sm_mhz": 1500.0, "sm_max_mhz": 1965.0, "reasons": ["hw_slowdown"] if hot else [], "samples": 3}

    mp.setattr(bench, "ClockSampler", ThrottledOnce)

bench.main()
print(f"rank {os.environ.get('RANK', '0')} finished", file=sys.stderr, flush=True)
