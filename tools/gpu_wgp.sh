#!/bin/bash
tag=${1:-wgp}
mkdir -p gpurun_out
for d in 0 1 2 3; do echo "== AGCN_WG_DEBUG=$d fp32"; AGCN_WG_DEBUG=$d timeout 300 python tools/bench_stage.py wgrad_tconv wgrad_proj_c64 wgrad_proj_c256; done > gpurun_out/${tag}_dbg_fp32.log 2>&1; cat gpurun_out/${tag}_dbg_fp32.log
