#!/bin/bash
mkdir -p gpurun_out
cp fusion_gcn_b200/libagcn_b200.so /tmp/cur.so
( echo "== previous build"; cp fusion_gcn_b200/prev_build.bin fusion_gcn_b200/libagcn_b200.so; timeout 300 python tools/bench_stage.py conv_tconv_c64 conv_tconv_c128
  echo "== current build, dual"; cp /tmp/cur.so fusion_gcn_b200/libagcn_b200.so; timeout 300 python tools/bench_stage.py conv_tconv_c64 conv_tconv_c128
  echo "== current build, single"; AGCN_TC2_NO_DUAL=1 timeout 300 python tools/bench_stage.py conv_tconv_c64 conv_tconv_c128 ) > gpurun_out/a16_stage.log 2>&1; cat gpurun_out/a16_stage.log
