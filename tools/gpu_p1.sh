#!/bin/bash
# limiter probes (wrong results, timing only): conv_tc2 / wgrad debug bits
mkdir -p gpurun_out
( for d in 0 1 2 3 4 7; do echo "== AGCN_CONV_DEBUG=$d fp32"; AGCN_CONV_DEBUG=$d timeout 200 python tools/bench_stage.py conv_tconv conv_proj_c256 conv_dproj_c64; done
  for d in 0 1 2 3; do echo "== AGCN_WG_DEBUG=$d fp32"; AGCN_WG_DEBUG=$d timeout 200 python tools/bench_stage.py wgrad_tconv wgrad_proj_c64; done
  echo "== tf32"; timeout 200 python tools/bench_stage.py conv_tconv wgrad_tconv conv_proj_c256 conv_dproj_c64 wgrad_proj_c64 --tf32 ) > gpurun_out/p1_probe.log 2>&1; cat gpurun_out/p1_probe.log
