#!/bin/bash
tag=${1:-a8}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stages.py -m gpu -q -x -k "conv or tf32_tensor" 2>&1 | tail -12
( echo "== tma store"; timeout 300 python tools/bench_stage.py conv_emb conv_proj conv_dproj convstats_proj res_dgrad
  echo "== old epilogue"; AGCN_TC2_NO_TMA_STORE=1 timeout 300 python tools/bench_stage.py conv_emb conv_proj conv_dproj convstats_proj res_dgrad
  echo "== tma store tf32"; timeout 300 python tools/bench_stage.py conv_emb conv_proj conv_dproj convstats_proj res_dgrad --tf32
  echo "== old epilogue tf32"; AGCN_TC2_NO_TMA_STORE=1 timeout 300 python tools/bench_stage.py conv_emb conv_proj conv_dproj convstats_proj res_dgrad --tf32 ) > gpurun_out/${tag}_stage.log 2>&1; cat gpurun_out/${tag}_stage.log
