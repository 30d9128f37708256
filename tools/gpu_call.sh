#!/usr/bin/env bash
# Round-2 GPU payloads (one gpurun call each).  Usage: bash tools/gpu_call.sh <name>
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bench_line() { python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('%.1f img/s  %.3f ms/step  e2e %.1f  launches/step %s roof %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d.get('gpu_launches_per_step'), (d.get('roofline') or {}).get('frac') or 0))"; }
case "${1}" in
  second)  # new default build (det + wide + 1x1): full suite, timeline, batch scaling, memcheck on the tiny model
    timeout 600 python -m pytest tests -m gpu -q --tb=short --timeout 300 > gpurun_out/r2_pytest2.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2_pytest2.log
    timeout 300 python tools/timeline.py > gpurun_out/r2_timeline_b32.txt 2>gpurun_out/r2_timeline_b32.err; head -50 gpurun_out/r2_timeline_b32.txt; tail -3 gpurun_out/r2_timeline_b32.err
    for b in 32 64; do printf "batch %d: " $b; timeout 300 python bench.py --batch $b --steps 10 --warmup 3 --skip-cpu-baseline 2>gpurun_out/r2_bench_b$b.err | tee gpurun_out/r2_bench_b$b.json | bench_line; done
    STEPS=2 timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python tools/repro_interleaved.py > gpurun_out/r2_memcheck_tiny.log 2>&1; echo "memcheck rc=$?"
    grep -E "ERROR SUMMARY|Invalid|step " gpurun_out/r2_memcheck_tiny.log | head -20 ;;
  third)  # side-stream weight gradients + batched alpha finish + drop-path: suite, A/B bench, timeline, ATen attribution
    timeout 600 python -m pytest tests -m gpu -q --tb=short --timeout 300 > gpurun_out/r2_pytest3.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2_pytest3.log
    for sw in 1 0; do printf "SGB_SIDE_WGRAD=%d: " $sw; SGB_SIDE_WGRAD=$sw timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline 2>gpurun_out/r2_bench_side$sw.err | tee gpurun_out/r2_bench_side$sw.json | bench_line; done
    timeout 300 python tools/timeline.py > gpurun_out/r2_timeline_side.txt 2>gpurun_out/r2_timeline_side.err; head -64 gpurun_out/r2_timeline_side.txt; tail -3 gpurun_out/r2_timeline_side.err
    timeout 300 python tools/timeline.py --aten --top 30 > gpurun_out/r2_aten_sites.txt 2>gpurun_out/r2_aten_sites.err; cat gpurun_out/r2_aten_sites.txt; tail -3 gpurun_out/r2_aten_sites.err ;;
  fourth)  # dy-slice backward, drop-path, eval fold cache, bench --config 2..5
    timeout 700 python -m pytest tests -m gpu -q --tb=short --timeout 300 > gpurun_out/r2_pytest4.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2_pytest4.log
    for c in 2 3 4 5; do printf "config %d: " $c; timeout 400 python bench.py --config $c --steps 10 --warmup 3 --skip-cpu-baseline 2>gpurun_out/r2_bench_c$c.err | tee gpurun_out/r2_bench_c$c.json | bench_line; tail -2 gpurun_out/r2_bench_c$c.err; done ;;
  fifth)  # patch stem, smem max-pool, linear-slot fix (config 4 capture), batched predict post-processing; memory-kernel table + ncu
    timeout 700 python -m pytest tests -m gpu -q --tb=short --timeout 300 > gpurun_out/r2_pytest5.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2_pytest5.log
    for c in 2 4 5; do printf "config %d: " $c; timeout 400 python bench.py --config $c --steps 10 --warmup 3 --skip-cpu-baseline 2>gpurun_out/r2_bench5_c$c.err | tee gpurun_out/r2_bench5_c$c.json | bench_line; tail -2 gpurun_out/r2_bench5_c$c.err; done
    printf "config 2, SGB_STEM_PATCHES=0: "; SGB_STEM_PATCHES=0 timeout 400 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline 2>/dev/null | bench_line
    timeout 300 python tools/mem_kernels.py > gpurun_out/r2_mem_kernels.txt 2>gpurun_out/r2_mem_kernels.err; cat gpurun_out/r2_mem_kernels.txt; tail -3 gpurun_out/r2_mem_kernels.err
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:"nms_kernel|loss_kernel|tal_topk_kernel|tal_decode_kernel" -c 8 -o gpurun_out/r2_full_mem -f python tools/mem_kernels.py --reps 1 > gpurun_out/r2_ncu_full_mem.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r2_ncu_full_mem.log
    timeout 300 python tools/timeline.py > gpurun_out/r2_timeline5.txt 2>gpurun_out/r2_timeline5.err; head -30 gpurun_out/r2_timeline5.txt ;;
  sixth)  # cooperative fused backward passes (A/B), smem-staged stem gather, ncu: nms + graph-step launch list + conv kernels
    timeout 700 python -m pytest tests -m gpu -q --tb=short --timeout 300 > gpurun_out/r2_pytest6.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2_pytest6.log
    for fb in 1 0; do printf "SGB_FUSED_BWD=%d: " $fb; SGB_FUSED_BWD=$fb timeout 400 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline 2>gpurun_out/r2_bench6_fb$fb.err | tee gpurun_out/r2_bench6_fb$fb.json | bench_line; tail -2 gpurun_out/r2_bench6_fb$fb.err; done
    timeout 300 python tools/timeline.py > gpurun_out/r2_timeline6.txt 2>gpurun_out/r2_timeline6.err; head -24 gpurun_out/r2_timeline6.txt
    timeout 400 ncu --set full --clock-control none --import-source on -k regex:"nms_kernel" -c 4 -o gpurun_out/r2_full_nms -f python tools/mem_kernels.py --reps 1 > gpurun_out/r2_ncu_full_nms.log 2>&1; echo "ncu nms rc=$?"
    SGB_PROFILER_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv       --log-file gpurun_out/r2_launches_graph_step.csv python bench.py --steps 1 --warmup 3 --skip-cpu-baseline > gpurun_out/r2_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
    for kn in conv_umma_kernel wgrad_umma_kernel; do
      timeout 400 ncu --set full --clock-control none --import-source on -k regex:$kn --launch-skip 20 --launch-count 3 -o gpurun_out/r2_full_$kn -f         python bench.py --steps 1 --warmup 1 --no-graph --skip-cpu-baseline > gpurun_out/r2_ncu_full_$kn.log 2>&1; echo "ncu $kn rc=$?"
    done ;;
  seventh)  # state of HEAD after the container was re-created: suite, fused-backward A/B, timeline with the per-launch dump
    timeout 700 python -m pytest tests -m gpu -q --tb=short --timeout 300 > gpurun_out/r2_pytest7.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2_pytest7.log
    for fb in 1 0; do printf "SGB_FUSED_BWD=%d: " $fb; SGB_FUSED_BWD=$fb timeout 400 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline 2>gpurun_out/r2_bench7_fb$fb.err | tee gpurun_out/r2_bench7_fb$fb.json | bench_line; tail -2 gpurun_out/r2_bench7_fb$fb.err; done
    timeout 300 python tools/timeline.py --dump gpurun_out/r2_timeline7_launches.txt > gpurun_out/r2_timeline7.txt 2>gpurun_out/r2_timeline7.err; head -40 gpurun_out/r2_timeline7.txt; tail -3 gpurun_out/r2_timeline7.err ;;
  eighth)  # cp.async ring in the per-channel passes, constant-folded stem gather, chunked layout kernels: suite, bench, timeline
    timeout 700 python -m pytest tests -m gpu -q --tb=short --timeout 300 -x > gpurun_out/r2_pytest8.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2_pytest8.log
    for fb in 1 0; do printf "SGB_FUSED_BWD=%d: " $fb; SGB_FUSED_BWD=$fb timeout 400 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline 2>gpurun_out/r2_bench8_fb$fb.err | tee gpurun_out/r2_bench8_fb$fb.json | bench_line; tail -2 gpurun_out/r2_bench8_fb$fb.err; done
    timeout 300 python tools/timeline.py --dump gpurun_out/r2_timeline8_launches.txt > gpurun_out/r2_timeline8.txt 2>gpurun_out/r2_timeline8.err; head -40 gpurun_out/r2_timeline8.txt; tail -3 gpurun_out/r2_timeline8.err ;;
  table)  # per-shape convolution table (eager, no overlap)
    timeout 300 python tools/conv_table.py > gpurun_out/r2_conv_table.txt 2>gpurun_out/r2_conv_table.err; cat gpurun_out/r2_conv_table.txt; tail -3 gpurun_out/r2_conv_table.err ;;
  ninth)  # shared input gradients (A/B), QARepVGG fold by map size
    timeout 700 python -m pytest tests -m gpu -q --tb=short --timeout 300 -x > gpurun_out/r2_pytest9.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_pytest9.log
    for sg in 1 0; do printf "SGB_SHARE_GRADS=%d: " $sg; SGB_SHARE_GRADS=$sg timeout 400 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline 2>gpurun_out/r2_bench9_sg$sg.err | tee gpurun_out/r2_bench9_sg$sg.json | bench_line; tail -2 gpurun_out/r2_bench9_sg$sg.err; done
    for mp in 12800 51200 204800 0; do printf "SGB_QAREP_FOLD=1 MAXPIX=%d: " $mp; SGB_QAREP_FOLD=1 SGB_QAREP_FOLD_MAXPIX=$mp timeout 400 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline 2>gpurun_out/r2_bench9_fold$mp.err | tee gpurun_out/r2_bench9_fold$mp.json | bench_line; tail -2 gpurun_out/r2_bench9_fold$mp.err; done
    timeout 300 python tools/timeline.py --dump gpurun_out/r2_timeline9_launches.txt > gpurun_out/r2_timeline9.txt 2>gpurun_out/r2_timeline9.err; head -30 gpurun_out/r2_timeline9.txt; tail -3 gpurun_out/r2_timeline9.err
    timeout 300 python tools/mem_kernels.py > gpurun_out/r2_mem_kernels9.txt 2>gpurun_out/r2_mem_kernels9.err; cat gpurun_out/r2_mem_kernels9.txt; tail -3 gpurun_out/r2_mem_kernels9.err ;;
  tenth)  # statistics inside the BatchNorm launch for wide layers, CSP conv1+conv2 as one GEMM, deferred shortcut gradients: suite subset, A/B, timeline
    timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_modules_gpu.py tests/test_trainer_gpu.py -m gpu -q --tb=short --timeout 300 -x > gpurun_out/r2_pytest10.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2_pytest10.log
    for cfg in "1 1 1" "0 1 1" "1 0 1" "1 1 0" "0 0 0"; do set -- $cfg; printf "STATS_IN_BN=%s DUAL_CONV=%s DEFER_SHORTCUT=%s: " $1 $2 $3
      SGB_STATS_IN_BN=$1 SGB_DUAL_CONV=$2 SGB_DEFER_SHORTCUT=$3 timeout 400 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline 2>gpurun_out/r2_bench10_$1$2$3.err | tee gpurun_out/r2_bench10_$1$2$3.json | bench_line; tail -2 gpurun_out/r2_bench10_$1$2$3.err; done
    timeout 300 python tools/timeline.py --dump gpurun_out/r2_timeline11_launches.txt > gpurun_out/r2_timeline11.txt 2>gpurun_out/r2_timeline11.err; head -45 gpurun_out/r2_timeline11.txt; tail -3 gpurun_out/r2_timeline11.err ;;
  eleventh)  # im2col kernels: 64-channel boxes with zero-filled tails, wide-N weight-gradient MMAs: kernel suite, A/B, conv table
    timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_modules_gpu.py -m gpu -q --tb=short --timeout 300 -x > gpurun_out/r2_pytest11.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2_pytest11.log
    for cfg in "1 1" "0 1" "1 0"; do set -- $cfg; printf "UMMA_KC_PAD=%s WGRAD_WIDE_N=%s: " $1 $2
      SGB_UMMA_KC_PAD=$1 SGB_WGRAD_WIDE_N=$2 timeout 400 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline 2>gpurun_out/r2_bench11_$1$2.err | tee gpurun_out/r2_bench11_$1$2.json | bench_line; tail -2 gpurun_out/r2_bench11_$1$2.err; done
    timeout 300 python tools/conv_table.py --top 150 > gpurun_out/r2_conv_table11.txt 2>gpurun_out/r2_conv_table11.err; grep -E "umma" gpurun_out/r2_conv_table11.txt | head -70; tail -3 gpurun_out/r2_conv_table11.err ;;
  twelfth)  # weight-gradient MMAs spanning taps, head cls/reg first convs as one GEMM
    timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_modules_gpu.py tests/test_trainer_gpu.py -m gpu -q --tb=short --timeout 300 -x > gpurun_out/r2_pytest12.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest12.log
    for wn in 1 0; do printf "WGRAD_WIDE_N=%s: " $wn
      SGB_WGRAD_WIDE_N=$wn timeout 400 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline 2>gpurun_out/r2_bench12_$wn.err | tee gpurun_out/r2_bench12_$wn.json | bench_line; tail -2 gpurun_out/r2_bench12_$wn.err; done
    SHAPES="32,48,320,320,96,3,2;32,96,160,160,192,3,2;32,96,160,160,96,3,2;32,192,80,80,384,3,2;32,96,160,160,64,1,1" timeout 120 python tools/conv_microbench.py wgrad
    timeout 300 python tools/timeline.py > gpurun_out/r2_timeline12.txt 2>gpurun_out/r2_timeline12.err; head -30 gpurun_out/r2_timeline12.txt; tail -3 gpurun_out/r2_timeline12.err ;;
  thirteenth)  # (historical: both experiments measured slower / neutral and were reverted -- SGB_UMMA_MSUB and SGB_OVERLAP_PREPARE no longer exist) two M tiles per filter tile in the im2col kernel, filter re-layout next to the stem gather
    timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_modules_gpu.py tests/test_trainer_gpu.py -m gpu -q --tb=short --timeout 300 -x > gpurun_out/r2_pytest13.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest13.log
    for cfg in "1 1" "0 1" "1 0"; do set -- $cfg; printf "UMMA_MSUB=%s OVERLAP_PREPARE=%s: " $1 $2
      SGB_UMMA_MSUB=$1 SGB_OVERLAP_PREPARE=$2 timeout 400 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline 2>gpurun_out/r2_bench13_$1$2.err | tee gpurun_out/r2_bench13_$1$2.json | bench_line; tail -2 gpurun_out/r2_bench13_$1$2.err; done
    for mode in fprop dgrad; do SHAPES="32,48,320,320,96,3,2;32,96,160,160,192,3,2;32,96,160,160,96,3,2;32,192,80,80,384,3,2;32,48,320,320,96,1,2" timeout 120 python tools/conv_microbench.py $mode; done ;;
  fourteenth)  # (historical: SGB_OVERLAP_PREPARE no longer exists) filter re-layout next to the stem gather (A/B) after reverting the two-M-tile experiment (slower: 1840 -> 1748 img/s)
    timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short --timeout 300 -x -k "conv_fprop" > gpurun_out/r2_pytest14.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest14.log
    for cfg in 1 0 1 0; do printf "OVERLAP_PREPARE=%s: " $cfg
      SGB_OVERLAP_PREPARE=$cfg timeout 400 python bench.py --steps 30 --warmup 5 --skip-cpu-baseline 2>gpurun_out/r2_bench14_$cfg.err | tee gpurun_out/r2_bench14_$cfg.json | bench_line; tail -2 gpurun_out/r2_bench14_$cfg.err; done ;;
  evidence_a)  # whole GPU suite + the four configuration bench lines
    timeout 900 python -m pytest tests -m gpu -q --tb=short --timeout 300 > gpurun_out/r2_pytest_final.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_pytest_final.log
    printf "config 2 (default command): "; timeout 500 python bench.py 2>gpurun_out/r2_final_c2.err | tee gpurun_out/r2_final_c2.json | bench_line; tail -2 gpurun_out/r2_final_c2.err
    for c in 3 4 5; do printf "config %d: " $c; timeout 400 python bench.py --config $c --steps 10 --warmup 3 --skip-cpu-baseline 2>gpurun_out/r2_final_c$c.err | tee gpurun_out/r2_final_c$c.json | bench_line; tail -2 gpurun_out/r2_final_c$c.err; done ;;
  evidence_b)  # ncu: launch list of ONE graph step with DRAM traffic, --set full of the top kernels (eager step)
    SGB_PROFILER_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv       --log-file gpurun_out/r2_launches_graph_step.csv python bench.py --steps 1 --warmup 3 --skip-cpu-baseline > gpurun_out/r2_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
    python tools/ncu_summary.py gpurun_out/r2_launches_graph_step.csv 60 > gpurun_out/r2_launches_graph_step.txt 2>&1; head -30 gpurun_out/r2_launches_graph_step.txt
    for kn in wgrad_umma_kernel conv_umma_kernel chan_fused_kernel conv3x3_halo_kernel wgrad3x3_halo_kernel; do
      timeout 400 ncu --set full --clock-control none --import-source on -k regex:$kn --launch-skip 12 --launch-count 4 -o gpurun_out/r2_full_$kn -f         python bench.py --steps 1 --warmup 1 --no-graph --skip-cpu-baseline > gpurun_out/r2_ncu_full_$kn.log 2>&1; echo "ncu $kn rc=$?"
      ncu -i gpurun_out/r2_full_$kn.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_pick.py > gpurun_out/r2_full_$kn.txt 2>&1; head -20 gpurun_out/r2_full_$kn.txt
    done ;;
  fifteenth)  # ResNet 7x7 stem as a 1x1 GEMM over gathered patches
    timeout 600 python -m pytest tests/test_modules_gpu.py tests/test_kernels_gpu.py -m gpu -q --tb=short --timeout 300 -x -k "stem or resnet or patch" > gpurun_out/r2_pytest15.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest15.log
    timeout 600 python -m pytest tests/test_zz_pose_train_gpu.py -m gpu -q --tb=short --timeout 300 -x -k "resnet" > gpurun_out/r2_pytest15b.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest15b.log
    for sp in 1 0; do printf "config 4 STEM_PATCHES=%s: " $sp; SGB_STEM_PATCHES=$sp timeout 400 python bench.py --config 4 --steps 10 --warmup 3 --skip-cpu-baseline 2>gpurun_out/r2_bench15_$sp.err | tee gpurun_out/r2_bench15_$sp.json | bench_line; tail -2 gpurun_out/r2_bench15_$sp.err; done
    timeout 400 python tools/conv_table.py --config 4 --all --top 24 > gpurun_out/r2_conv_table_resnet50_b.txt 2>gpurun_out/r2_conv_table_resnet50_b.err; tail -3 gpurun_out/r2_conv_table_resnet50_b.err; cat gpurun_out/r2_conv_table_resnet50_b.txt ;;
  dp)  # N GPUs (gpurun --gpus N): split graphs around an eager all-reduce vs NCCL captured inside one graph
    N=${2:-2}
    for ig in 0 1; do
      printf "N=%d SGB_NCCL_IN_GRAPH=%d: " $N $ig
      SGB_NCCL_IN_GRAPH=$ig timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520 + ig)) bench.py --gpus $N --steps 30 --warmup 5 --skip-cpu-baseline \
        > gpurun_out/r2_bench_dp${N}_ig$ig.json 2> gpurun_out/r2_bench_dp${N}_ig$ig.err; echo "rc=$?"; tail -2 gpurun_out/r2_bench_dp${N}_ig$ig.err; bench_line < gpurun_out/r2_bench_dp${N}_ig$ig.json
    done ;;
  tenth)  # shared grads off by default; fold / fused-forward A/B; timeline of the folded step
    timeout 700 python -m pytest tests -m gpu -q --tb=short --timeout 300 -x > gpurun_out/r2_pytest10.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_pytest10.log
    for v in "SGB_QAREP_FOLD=0 SGB_FUSED_FWD=1" "SGB_QAREP_FOLD=1 SGB_FUSED_FWD=1" "SGB_QAREP_FOLD=0 SGB_FUSED_FWD=0" "SGB_QAREP_FOLD=1 SGB_FUSED_FWD=0"; do
      printf "%s: " "$v"; env $v timeout 400 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline 2>gpurun_out/r2_bench10.err | tee "gpurun_out/r2_bench10_$(echo $v | tr ' =' '__').json" | bench_line; tail -2 gpurun_out/r2_bench10.err
    done
    SGB_QAREP_FOLD=1 timeout 300 python tools/timeline.py --dump gpurun_out/r2_timeline10_launches.txt > gpurun_out/r2_timeline10.txt 2>gpurun_out/r2_timeline10.err; head -36 gpurun_out/r2_timeline10.txt; tail -3 gpurun_out/r2_timeline10.err
    SGB_QAREP_FOLD=1 timeout 300 python tools/conv_table.py --all --top 60 > gpurun_out/r2_conv_table10.txt 2>gpurun_out/r2_conv_table10.err; tail -3 gpurun_out/r2_conv_table10.err ;;
  eleventh)  # split NMS (front / IoU matrix / back), batched folded plumbing; engine-attributed conv table
    timeout 700 python -m pytest tests -m gpu -q --tb=short --timeout 300 -x > gpurun_out/r2_pytest11.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_pytest11.log
    for v in "SGB_QAREP_FOLD=0" "SGB_QAREP_FOLD=1"; do
      printf "%s: " "$v"; env $v timeout 400 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline 2>gpurun_out/r2_bench11.err | tee "gpurun_out/r2_bench11_$(echo $v | tr ' =' '__').json" | bench_line; tail -2 gpurun_out/r2_bench11.err
    done
    timeout 300 python tools/mem_kernels.py > gpurun_out/r2_mem_kernels11.txt 2>gpurun_out/r2_mem_kernels11.err; cat gpurun_out/r2_mem_kernels11.txt; tail -3 gpurun_out/r2_mem_kernels11.err
    timeout 300 python tools/conv_table.py --top 300 > gpurun_out/r2_conv_table11.txt 2>gpurun_out/r2_conv_table11.err; head -8 gpurun_out/r2_conv_table11.txt; tail -3 gpurun_out/r2_conv_table11.err ;;
  twelfth)  # in-place batched folded filters; NMS kernel breakdown
    timeout 700 python -m pytest tests -m gpu -q --tb=short --timeout 300 -x > gpurun_out/r2_pytest12.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_pytest12.log
    for v in "SGB_QAREP_FOLD=0" "SGB_QAREP_FOLD=1"; do
      printf "%s: " "$v"; env $v timeout 400 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline 2>gpurun_out/r2_bench12.err | tee "gpurun_out/r2_bench12_$(echo $v | tr ' =' '__').json" | bench_line; tail -2 gpurun_out/r2_bench12.err
    done
    timeout 300 python tools/mem_kernels.py --breakdown > gpurun_out/r2_mem_kernels12.txt 2>gpurun_out/r2_mem_kernels12.err; cat gpurun_out/r2_mem_kernels12.txt; tail -3 gpurun_out/r2_mem_kernels12.err
    SGB_QAREP_FOLD=1 timeout 300 python tools/timeline.py --dump gpurun_out/r2_timeline12_launches.txt > gpurun_out/r2_timeline12.txt 2>gpurun_out/r2_timeline12.err; head -40 gpurun_out/r2_timeline12.txt; tail -3 gpurun_out/r2_timeline12.err ;;
  thirteenth)  # fold default; NMS atomics / mapping fixes
    timeout 700 python -m pytest tests -m gpu -q --tb=short --timeout 300 -x > gpurun_out/r2_pytest13.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_pytest13.log
    printf "default: "; timeout 400 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline 2>gpurun_out/r2_bench13.err | tee gpurun_out/r2_bench13.json | bench_line; tail -2 gpurun_out/r2_bench13.err
    timeout 300 python tools/mem_kernels.py --breakdown > gpurun_out/r2_mem_kernels13.txt 2>gpurun_out/r2_mem_kernels13.err; cat gpurun_out/r2_mem_kernels13.txt; tail -3 gpurun_out/r2_mem_kernels13.err
    for c in 3 4 5; do printf "config %d: " $c; timeout 400 python bench.py --config $c --steps 10 --warmup 3 --skip-cpu-baseline 2>gpurun_out/r2_bench13_c$c.err | tee gpurun_out/r2_bench13_c$c.json | bench_line; tail -2 gpurun_out/r2_bench13_c$c.err; done ;;
  ncu_nms)  # ncu --set full of the NMS kernels (detection callback at benchmark size)
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:"nms_" -c 6 -o gpurun_out/r2_full_nms -f python tools/mem_kernels.py --reps 1 > gpurun_out/r2_ncu_full_nms.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r2_ncu_full_nms.log
    ls -la gpurun_out/*.ncu-rep ;;
  multi)  # N GPUs (gpurun --gpus N): the BASELINE bench + the in-situ timeline of every rank (all-reduce duration, skew)
    N=${2:-2}
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 --skip-cpu-baseline \
      > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_n$N.err; bench_line < gpurun_out/r2_bench_n$N.json
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 tools/timeline.py --steps 4 --top 12 > gpurun_out/r2_timeline_n$N.txt 2> gpurun_out/r2_timeline_n$N.err
    echo "timeline rc=$?"; head -16 gpurun_out/r2_timeline_n$N.txt; grep -h "nccl\|wall" gpurun_out/timeline_rank*.txt | head -40 ;;
esac
