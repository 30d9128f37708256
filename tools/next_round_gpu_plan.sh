#!/usr/bin/env bash
# First GPU calls of the next round, cheapest / most informative first (each line is one `gpurun` payload; wrap in `timeout`).
#   1. bash tools/next_round_gpu_plan.sh verify      -> every GPU test incl. the ones written after round 1's budget ran out
#   1b. bash tools/next_round_gpu_plan.sh sanitize   -> compute-sanitizer memcheck over the small fixtures of the never-run kernels
#   2. bash tools/next_round_gpu_plan.sh repro       -> the interleaved-model anomaly under compute-sanitizer
#   2b. bash tools/next_round_gpu_plan.sh determinism -> same scenario on the -DSGB_DETERMINISTIC_STATS build (prebuilt HERE by
#       `SGB_OUT=$PWD/super_gradients_b200/libsgb200_det.so SGB_OBJ=$PWD/super_gradients_b200/csrc/obj_det bash super_gradients_b200/csrc/build.sh -DSGB_DETERMINISTIC_STATS`)
#   2c. bash tools/next_round_gpu_plan.sh pdl -> programmatic dependent launch build (-DSGB_PDL, prebuilt HERE the same way into
#       super_gradients_b200/libsgb200_pdl.so): kernel / module / trainer suites, then default vs PDL bench back to back
#   2d. bash tools/next_round_gpu_plan.sh variants -> default / _det / _pdl / _wide (-DSGB_UMMA_WIDE_STORE: 256-bit stores in the
#       im2col kernels' fast epilogue) / _exp (all flags) benched back to back
#       _1x1 (-DSGB_HALO_1X1: 1x1 stride-1 convolutions on the halo kernel's pipeline with a plain 16 x 16 tile; run
#       `SGB200_LIB=.../libsgb200_1x1.so pytest tests/test_kernels_gpu.py tests/test_zz_pose_train_gpu.py -m gpu --runxfail -k conv` first)
#   3. bash tools/next_round_gpu_plan.sh twogpu      -> 2-GPU bench, hard 150 s limit (run with `gpurun --gpus 2`)
#   4. bash tools/next_round_gpu_plan.sh profile     -> ncu launch list of one graph step + layer profile
# `bash tools/build_variants.sh` (here, no GPU needed) builds all experiment libraries; rerun it after any change to include/sgb200.h.
# NOTE: the experiment libraries (libsgb200_{det,pdl,wide,exp}.so) are listed in .gpurunignore so that routine calls stay small:
# comment those lines out before `determinism`, `pdl` or `variants`.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
case "${1:-verify}" in
  verify)
    timeout 500 python -m pytest tests -m gpu -q --tb=short --timeout 300 2>&1 | tail -25
    # the tests marked xfail (everything written without hardware: pose rows, YoloX NMS, pre-processing, DetectionMetrics matching, ...) with their real outcome
    # one process per group: an illegal address in one new kernel must not take the other groups' results with it
    for grp in "pose_loss or pose_assigner" "tiny_yolo_nas_pose" "yolox" "preprocessing or raw_images" "folded" "conv_1x1" "adjoint" "split_graph" \
               "detection_matching" "atss" "focal" "vs_reference and not golden"; do
      echo "== -k '$grp'"; timeout 400 python -m pytest tests/test_zz_pose_train_gpu.py -m gpu -q --tb=line --runxfail -k "$grp" 2>&1 | tail -6
    done
    timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
    timeout 300 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_verify.json 2> gpurun_out/bench_verify.err
    tail -2 gpurun_out/bench_verify.err; cut -c1-400 gpurun_out/bench_verify.json ;;
  sanitize)  # memcheck over the kernels that have never run (small fixtures only: ~40x slowdown)
    for grp in "detection_matching_kernel_vs" "atss_assigner_and_static" "focal" "preprocessing" "pose_loss_kernels_match_the_reference" "yolox"; do
      echo "== memcheck -k '$grp'"
      timeout 600 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_zz_pose_train_gpu.py -m gpu -q --tb=line --runxfail -k "$grp" > gpurun_out/sanitize_${grp// /_}.log 2>&1
      grep -E "ERROR SUMMARY|Invalid|passed|failed" gpurun_out/sanitize_${grp// /_}.log | tail -4
    done ;;
  repro)
    STEPS=4 timeout 120 python tools/repro_interleaved.py 2>&1 | tail -6
    for tool in memcheck initcheck; do
      echo "== compute-sanitizer $tool"
      STEPS=3 timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/repro_interleaved.py > gpurun_out/sanitizer_$tool.log 2>&1
      grep -E "ERROR SUMMARY|Invalid|Uninitialized|at 0x|by thread" gpurun_out/sanitizer_$tool.log | head -30
    done ;;
  determinism)
    echo "== default build"; STEPS=6 timeout 120 python tools/repro_interleaved.py 2>&1 | tail -8
    echo "== deterministic per-warp statistics slots"
    SGB200_LIB=$PWD/super_gradients_b200/libsgb200_det.so STEPS=6 timeout 120 python tools/repro_interleaved.py 2>&1 | tail -8
    SGB200_LIB=$PWD/super_gradients_b200/libsgb200_det.so timeout 400 python -m pytest tests/test_kernels_gpu.py tests/test_modules_gpu.py -m gpu -q --tb=short 2>&1 | tail -5
    SGB200_LIB=$PWD/super_gradients_b200/libsgb200_det.so timeout 300 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline 2>/dev/null | cut -c1-300 ;;
  variants)  # every prepared experiment, one bench each (each library is a full build with one -D flag; exp = all of them)
    for v in "" _det _pdl _wide _1x1 _exp; do
      lib=$PWD/super_gradients_b200/libsgb200$v.so; [[ -f $lib ]] || { echo "missing $lib (see the header of this script)"; continue; }
      printf "%-22s" "libsgb200$v.so"; SGB200_LIB=$lib timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline 2>gpurun_out/bench_variant$v.err | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('%.1f img/s  %.3f ms/step  e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
    done
    # Python-level experiment on the default library: QARepVGG 1x1 branch folded into the 3x3 convolution (functional.QAREP_FOLD)
    printf "%-22s" "SGB_QAREP_FOLD=1"; SGB_QAREP_FOLD=1 timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline 2>gpurun_out/bench_variant_fold.err | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('%.1f img/s  %.3f ms/step  e2e %.1f  launches/step %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches_per_step']))"
    # cost of the data-parallel capture layout (two graphs around an eager collective) measured on one GPU
    printf "%-22s" "SGB_SPLIT_GRAPH=1"; SGB_SPLIT_GRAPH=1 timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline 2>gpurun_out/bench_variant_split.err | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('%.1f img/s  %.3f ms/step' % (d['value'], d['ms_per_step']))" ;;
  pdl)
    SGB200_LIB=$PWD/super_gradients_b200/libsgb200_pdl.so timeout 500 python -m pytest tests/test_kernels_gpu.py tests/test_modules_gpu.py tests/test_trainer_gpu.py -m gpu -q --tb=short -x 2>&1 | tail -6
    for lib in libsgb200.so libsgb200_pdl.so; do
      echo "== $lib"; SGB200_LIB=$PWD/super_gradients_b200/$lib timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline 2>gpurun_out/bench_$lib.err | cut -c1-260
    done ;;
  twogpu)
    timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus 2 --steps 8 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
    echo "rc=$?"; tail -5 gpurun_out/bench_2gpu.err; cut -c1-500 gpurun_out/bench_2gpu.json ;;
  profile)
    SGB_PROFILER_RANGE=1 timeout 700 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
      --clock-control none --csv --log-file gpurun_out/launches_next.csv python bench.py --steps 1 --warmup 3 --skip-cpu-baseline > gpurun_out/ncu_next.log 2>&1
    timeout 200 python tools/layer_profile.py > gpurun_out/layers_next.txt 2>&1; head -40 gpurun_out/layers_next.txt
    # what limits the streaming BN / QARepVGG passes (2.3-3.3 TB/s of 6.6) and the im2col kernel on 1x1 layers: one full capture each
    for kn in "chan_kernel" "conv_umma_kernel" "wgrad_umma_kernel"; do
      timeout 400 ncu --set full --clock-control none --import-source on -k regex:$kn --launch-skip 12 --launch-count 2 -o gpurun_out/full_$kn -f \
        python bench.py --steps 1 --warmup 1 --no-graph --skip-cpu-baseline > gpurun_out/ncu_full_$kn.log 2>&1
      ncu -i gpurun_out/full_$kn.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_pick.py 2>/dev/null | head -30
    done ;;
esac
