#!/bin/bash
mkdir -p gpurun_out
( for i in 1 2; do timeout 300 python tools/bench_stage.py mix_score_bwd_c64 mix_fwd_c64; done
  echo "== tma score"; AGCN_MIX_TMA_STORE_SCORE=1 timeout 300 python tools/bench_stage.py mix_score_bwd_c64
  echo "== simt score"; AGCN_MIX_SCORE_SIMT=1 timeout 300 python tools/bench_stage.py mix_score_bwd_c64
  nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv ) > gpurun_out/a13.log 2>&1; cat gpurun_out/a13.log
