#!/usr/bin/env bash
# Builds the experiment libraries next to the default one (each a full build with one -D flag; _exp = all of them).  They are
# git-ignored and listed in .gpurunignore: comment those lines out before a gpurun call that needs them
# (tools/next_round_gpu_plan.sh determinism / pdl / variants).  Rebuild after any change to include/sgb200.h: lib.py binds every
# declared entry point at load time, so a stale variant fails to load.
set -euo pipefail
cd "$(dirname "$0")/.."
P=$PWD/super_gradients_b200
build() { SGB_OUT=$P/libsgb200_$1.so SGB_OBJ=$P/csrc/obj_$1 bash $P/csrc/build.sh "${@:2}" 2>&1 | tail -1; }
build det -DSGB_DETERMINISTIC_STATS
build pdl -DSGB_PDL
build wide -DSGB_UMMA_WIDE_STORE
build 1x1 -DSGB_HALO_1X1
build exp -DSGB_DETERMINISTIC_STATS -DSGB_PDL -DSGB_UMMA_WIDE_STORE -DSGB_HALO_1X1
