#!/usr/bin/env python
"""Are the kernels measured on hardware still the same machine code?  Rebuilds the .cu files of a past commit (default: the last
commit whose kernels ran on a B200) into a scratch directory and compares, function by function, the instruction stream of every
kernel that exists in both builds with the current objects under super_gradients_b200/csrc/obj (run csrc/build.sh first).
Names are compared with the translation-unit hash of anonymous namespaces removed; kernels that only exist now are listed.
Usage: python tools/sass_identity_check.py [commit]      (exit code 1 if a common kernel differs)"""
import hashlib
import os
import re
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VERIFIED = "2f1a0e8"  # round 1's last hardware-verified build
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def functions(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    res, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = re.sub(r"_GLOBAL__N__[0-9a-f]+_\d+_\w+?_cu_[0-9a-f]+", "ANON", m.group(1))
            res[cur] = []
        elif cur and re.match(r"^\s+/\*[0-9a-f]{4}\*/", line):
            res[cur].append(line)
    return {k: hashlib.md5("\n".join(v).encode()).hexdigest() for k, v in res.items()}


def main(commit):
    scratch = tempfile.mkdtemp(prefix="sass_check_", dir=os.path.join(ROOT, "gpurun_out") if os.path.isdir(os.path.join(ROOT, "gpurun_out")) else None)
    tar = subprocess.run(["git", "-C", ROOT, "archive", commit, "super_gradients_b200/csrc", "include"], capture_output=True, check=True).stdout
    subprocess.run(["tar", "-x", "-C", scratch], input=tar, check=True)
    src = os.path.join(scratch, "super_gradients_b200", "csrc")
    bad = 0
    procs = {}
    for f in sorted(os.listdir(src)):
        if f.endswith(".cu") and os.path.exists(os.path.join(ROOT, "super_gradients_b200", "csrc", "obj", f[:-3] + ".o")):
            o = os.path.join(scratch, f[:-3] + ".o")
            procs[f] = (o, subprocess.Popen(["nvcc", *FLAGS, "-I", os.path.join(scratch, "include"), "-I", src, "-c", os.path.join(src, f), "-o", o], stderr=subprocess.DEVNULL))
    for f, (o, p) in procs.items():
        p.wait()
        old, new = functions(o), functions(os.path.join(ROOT, "super_gradients_b200", "csrc", "obj", f[:-3] + ".o"))
        diff = [k for k in old if k in new and old[k] != new[k]]
        gone = [k for k in old if k not in new]
        added = [k for k in new if k not in old]
        print(f"{f:24s} {len(old) - len(diff) - len(gone):3d} identical, {len(diff)} changed, {len(gone)} removed, {len(added)} new")
        for k in diff + gone:
            print("    !", k[-90:])
        bad += len(diff) + len(gone)
    shutil.rmtree(scratch, ignore_errors=True)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1] if len(sys.argv) > 1 else VERIFIED))
