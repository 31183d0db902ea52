#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
( echo "== dual (2 MMAs)"; timeout 300 python tools/bench_stage.py conv_tconv_c64 conv_tconv_c128 convstats_tconv_c64
  echo "== AGCN_TC2_NO_DUAL"; AGCN_TC2_NO_DUAL=1 timeout 300 python tools/bench_stage.py conv_tconv_c64 ) > gpurun_out/a23_stage.log 2>&1; cat gpurun_out/a23_stage.log
