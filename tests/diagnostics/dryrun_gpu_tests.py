#!/usr/bin/env python
"""Runs the bodies of `-m gpu` test functions on the CPU stand-in of the kernel wrappers (tests/cpu_backend.py) with DEV = "cpu":
it cannot say anything about the kernels, but it executes the TEST code (fixture keys, shapes, helper names, tolerances against the
host build of the kernel arithmetic), so a test written without hardware does not waste its first GPU call on a typo.
Usage: python tests/diagnostics/dryrun_gpu_tests.py tests/test_zz_pose_train_gpu.py [name-substring-to-skip ...]"""
import importlib
import inspect
import itertools
import os
import pathlib
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch  # noqa: E402
from _pytest.monkeypatch import MonkeyPatch  # noqa: E402

import cpu_backend  # noqa: E402


def main(path, skip):
    mp = MonkeyPatch()
    cpu_backend.install_training(mp)
    mp.setattr(torch.cuda, "synchronize", lambda *a: None)
    mod = importlib.import_module(os.path.splitext(os.path.basename(path))[0])
    mod.DEV = "cpu"
    cache = {}

    def golden(name):
        if name not in cache:
            cache[name] = torch.load(os.path.join(ROOT, "tests", "golden", name + ".pt"), weights_only=False)
        return cache[name]

    ok = bad = 0
    for name, fn in inspect.getmembers(mod, inspect.isfunction):
        if not name.startswith("test_") or any(s in name for s in skip):
            continue
        marks = [m for m in getattr(fn, "pytestmark", []) if m.name == "parametrize"]
        names, values = [m.args[0] for m in marks], [list(m.args[1]) for m in marks]
        for combo in itertools.product(*values) if values else [()]:
            kwargs = {}
            for n, v in zip(names, combo):
                if "," in n:
                    kwargs.update(zip([x.strip() for x in n.split(",")], v))
                else:
                    kwargs[n] = v
            sig = inspect.signature(fn).parameters
            if "golden" in sig:
                kwargs["golden"] = golden
            if "tmp_path" in sig:
                kwargs["tmp_path"] = pathlib.Path(tempfile.mkdtemp())
            local_mp = MonkeyPatch()
            if "monkeypatch" in sig:
                kwargs["monkeypatch"] = local_mp
            try:
                fn(**kwargs)
                ok += 1
            except Exception as e:  # noqa: BLE001
                bad += 1
                print("FAIL", name, combo, type(e).__name__, str(e)[:300])
            finally:
                local_mp.undo()
    print(f"{ok} passed, {bad} failed")
    return bad


if __name__ == "__main__":
    sys.exit(1 if main(sys.argv[1], sys.argv[2:]) else 0)
