#!/bin/bash
mkdir -p gpurun_out
( for d in 0 1 2 3 4 7; do echo "== AGCN_CONV_DEBUG=$d fp32"; AGCN_CONV_DEBUG=$d timeout 200 python tools/bench_stage.py conv_emb_c64 conv_proj_c64 conv_dproj_c64 conv_proj_c256; done
  echo "== tf32"; timeout 200 python tools/bench_stage.py conv_emb_c64 conv_proj_c64 conv_dproj_c64 conv_proj_c256 --tf32 ) > gpurun_out/p2_probe.log 2>&1; cat gpurun_out/p2_probe.log
