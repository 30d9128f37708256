"""The C-ABI library loads and exports exactly the symbols include/sgb200.h declares (no compute calls: no GPU here)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "sgb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sgb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from super_gradients_b200 import lib

    names = declared_symbols()
    assert len(names) >= 35
    assert sorted(lib.exported_names()) == names, set(lib.exported_names()) ^ set(names)
    handle = lib.load()  # binds every symbol; AttributeError if one is missing from the .so
    for n in names:
        assert hasattr(handle, n)
    assert handle.sgb_version() >= 100


def test_product_path_fails_loudly_without_cuda():
    import pytest
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from super_gradients_b200.lib import SgbError
    from super_gradients_b200.modules import QARepVGGBlock
    from super_gradients_b200.training import models

    with pytest.raises(SgbError):
        QARepVGGBlock(8, 8)(torch.zeros(1, 8, 4, 4))
    with pytest.raises(SgbError):
        models.get("resnet18_cifar", num_classes=10)(torch.zeros(1, 3, 32, 32))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "super_gradients_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(d, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), os.path.join(d, f)
    # developer tools stay oracle-free too (anything that needs the checker lives under tests/)
    for f in os.listdir(os.path.join(ROOT, "tools")):
        if f.endswith(".py"):
            src = open(os.path.join(ROOT, "tools", f)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_ctypes_structs_match_the_c_header(tmp_path):
    """Every struct of include/sgb200.h, compiled by gcc, has the size and field offsets of its ctypes mirror in lib.py
    (a silent mismatch would hand the kernels garbage descriptors)."""
    import ctypes
    import re
    import shutil
    import subprocess

    from super_gradients_b200 import lib as L

    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    pairs = {"SgbConvDesc": L.ConvDesc, "SgbEpilogue": L.Epilogue, "SgbWeightItem": L.WeightItem, "SgbWgradItem": L.WgradItem, "SgbAlphaItem": L.AlphaItem,
             "SgbBnDesc": L.BnDesc, "SgbQarepDesc": L.QarepDesc, "SgbLossDesc": L.LossDesc, "SgbPoseLossDesc": L.PoseLossDesc, "SgbNmsDesc": L.NmsDesc, "SgbPreprocDesc": L.PreprocDesc, "SgbMatchDesc": L.MatchDesc}  # fmt: skip
    header = open(os.path.join(ROOT, "include", "sgb200.h")).read()
    assert set(re.findall(r"typedef struct (Sgb\w+)", header)) == set(pairs), "a header struct has no ctypes mirror (or vice versa)"
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "sgb200.h"', "int main(void) {"]
    for cname, cls in pairs.items():
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    for line in out.strip().splitlines():
        cname, field, value = line.split()
        cls = pairs[cname]
        expect = ctypes.sizeof(cls) if field == "size" else getattr(cls, field).offset
        assert int(value) == expect, (cname, field, int(value), expect)
