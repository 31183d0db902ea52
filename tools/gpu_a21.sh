#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stages.py -m gpu -q -x -k "conv or tf32_tensor" 2>&1 | tail -3
( echo "== resident"; timeout 300 python tools/bench_stage.py conv_emb conv_proj conv_dproj convstats_proj res_dgrad; timeout 300 python tools/bench_stage.py conv_emb conv_proj conv_dproj --tf32
  echo "== streamed (AGCN_TC2_NO_RESIDENT_W)"; AGCN_TC2_NO_RESIDENT_W=1 timeout 300 python tools/bench_stage.py conv_emb_c64 conv_proj_c64 conv_dproj_c64 conv_emb_c128; AGCN_TC2_NO_RESIDENT_W=1 timeout 300 python tools/bench_stage.py conv_emb_c64 conv_proj_c64 conv_dproj_c64 conv_emb_c128 conv_proj_c128 --tf32 ) > gpurun_out/a21_stage.log 2>&1; cat gpurun_out/a21_stage.log
