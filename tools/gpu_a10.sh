#!/bin/bash
mkdir -p gpurun_out
( echo "== default"; timeout 300 python tools/bench_stage.py conv_tconv; timeout 300 python tools/bench_stage.py conv_tconv --tf32
  echo "== AGCN_TC2_TMA_STORE_ALL"; AGCN_TC2_TMA_STORE_ALL=1 timeout 300 python tools/bench_stage.py conv_tconv;  AGCN_TC2_TMA_STORE_ALL=1 timeout 300 python tools/bench_stage.py conv_tconv --tf32
  AGCN_TC2_TMA_STORE_ALL=1 timeout 600 python -m pytest tests/test_gpu_stages.py -m gpu -q -x -k "conv or tf32_tensor" 2>&1 | tail -3 ) > gpurun_out/a10_stage.log 2>&1; cat gpurun_out/a10_stage.log
