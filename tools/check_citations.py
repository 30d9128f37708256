#!/usr/bin/env python
"""Every `path/to/file.py:LINE[-LINE]` citation in the sources, the C header and the documents must name lines that exist in the
reference tree (run in the build container, where /root/reference is mounted; a citation whose file cannot be resolved uniquely by
its path suffix is reported as ambiguous, not as an error).  Usage: python tools/check_citations.py [reference_root]"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
PAT = re.compile(r"([\w/\.\-]+\.(?:py|yaml)):(\d+)(?:-(\d+))?")


def main():
    if not os.path.isdir(REF):
        print("reference tree not present: nothing to check")
        return 0
    index = {}
    for d, _, files in os.walk(REF):
        for f in files:
            if f.endswith((".py", ".yaml")):
                index.setdefault(f, []).append(os.path.join(d, f))
    lengths, bad, n = {}, [], 0
    srcs = [os.path.join(ROOT, f) for f in ("DESIGN.md", "INTEGRATION.md", "README.md", "include/sgb200.h")]
    for top in ("super_gradients_b200", "oracle", "tests"):
        for d, _, files in os.walk(os.path.join(ROOT, top)):
            srcs += [os.path.join(d, f) for f in files if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".inc"))]
    for src in srcs:
        for ln, line in enumerate(open(src, errors="replace"), 1):
            for m in PAT.finditer(line):
                path, a, b = m.group(1), int(m.group(2)), int(m.group(3) or m.group(2))
                cands = [p for p in index.get(os.path.basename(path), []) if p.endswith("/" + path.lstrip("./")) or os.path.basename(path) == path]
                if len(cands) != 1:
                    continue  # own files (tests/..., csrc/...) or an ambiguous basename
                n += 1
                if cands[0] not in lengths:
                    lengths[cands[0]] = sum(1 for _ in open(cands[0], errors="replace"))
                if b > lengths[cands[0]] or a > b or a < 1:
                    bad.append(f"{os.path.relpath(src, ROOT)}:{ln}: {m.group(0)} but {os.path.relpath(cands[0], REF)} has {lengths[cands[0]]} lines")
    print(f"{n} citations resolved, {len(bad)} out of range")
    for b in bad:
        print("  ", b)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
