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
esac
