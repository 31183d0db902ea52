#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stages.py -m gpu -q -x -k "conv" 2>&1 | tail -4
( echo "== dual"; timeout 300 python tools/bench_stage.py conv_tconv_c64 convstats_tconv_c64
  echo "== single (AGCN_TC2_NO_DUAL)"; AGCN_TC2_NO_DUAL=1 timeout 300 python tools/bench_stage.py conv_tconv_c64 convstats_tconv_c64 ) > gpurun_out/a15_stage.log 2>&1; cat gpurun_out/a15_stage.log
