./tools/umma_probe > gpurun_out/probe.txt 2>&1; tail -1 gpurun_out/probe.txt
export SHAPES="32,32,160,160,32,3,1;32,64,80,80,64,3,1;32,96,160,160,96,1,1;32,96,40,40,96,3,1;32,48,80,80,48,3,1;32,192,20,20,192,3,1"
export NOSTATS=1
for cfg in "1 8" "1 12" "1 18" "2 8" "3 6" "2 12"; do set -- $cfg; echo "== CTAS=$1 STAGES=$2"; SGB_CTAS_PER_SM=$1 SGB_MAX_STAGES=$2 timeout 100 python tools/conv_microbench.py fprop 2>&1 | grep "^fprop"; done
