#!/bin/bash
tag=${1:-dbg}
mkdir -p gpurun_out
L=gpurun_out/${tag}_conv_dbg.log; : > $L
for v in new old; do
  for d in 0 1 2 4 7; do
    if [ $v = old ]; then export AGCN_TC2_NO1X1=1; else unset AGCN_TC2_NO1X1; fi
    echo "== 1x1 via $v kernel, AGCN_CONV_DEBUG=$d fp32" >> $L
    AGCN_CONV_DEBUG=$d timeout 300 python tools/bench_stage.py conv_emb_c64 conv_proj_c64 conv_dproj_c64 conv_tconv_c64 conv_proj_c256 >> $L 2>&1
  done
  for d in 0 2 4 6; do
    echo "== 1x1 via $v kernel, AGCN_CONV_DEBUG=$d tf32" >> $L
    AGCN_CONV_DEBUG=$d timeout 300 python tools/bench_stage.py conv_emb_c64 conv_proj_c64 conv_dproj_c64 conv_tconv_c64 conv_proj_c256 --tf32 >> $L 2>&1
  done
done
cat $L
unset AGCN_TC2_NO1X1
timeout 300 python tools/bench_stage.py mix gram > gpurun_out/${tag}_joint.log 2>&1; cat gpurun_out/${tag}_joint.log
timeout 300 python -m pytest tests/test_gpu_stages.py -m gpu -x -q -k "mix or gram" 2>&1 | tail -3
