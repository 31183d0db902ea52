#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stages.py -m gpu -q -x -k "mix" 2>&1 | tail -6
( echo "== tma out"; timeout 300 python tools/bench_stage.py mix_; timeout 300 python tools/bench_stage.py mix_ --tf32
  echo "== old"; AGCN_MIX_NO_TMA_STORE=1 timeout 300 python tools/bench_stage.py mix_c64 mix_fwd_c64 mix_bwd_c64 mix_score_bwd_c64; AGCN_MIX_NO_TMA_STORE=1 timeout 300 python tools/bench_stage.py mix_fwd_c64 mix_bwd_c64 mix_score_bwd_c64 --tf32 ) > gpurun_out/a11_stage.log 2>&1; cat gpurun_out/a11_stage.log
