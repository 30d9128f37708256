export SHAPES="32,48,320,320,96,3,2;32,96,160,160,192,3,2;32,96,160,160,96,3,2"
for mode in wgrad fprop dgrad; do
  for m in 0 1 4 8 12 5 13; do
    echo "== $mode SGB_DEBUG_SKIP=$m"; SGB_DEBUG_SKIP=$m timeout 120 python tools/conv_microbench.py $mode 2>&1 | grep -v "^sm100"
  done
done > gpurun_out/r2_s2_microbench.txt 2>&1
cat gpurun_out/r2_s2_microbench.txt
