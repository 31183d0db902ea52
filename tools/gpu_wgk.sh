#!/bin/bash
tag=${1:-wgk}
mkdir -p gpurun_out
for k in 24 32 40 48; do echo "== AGCN_WG_SPLIT_KB=$k fp32"; AGCN_WG_SPLIT_KB=$k timeout 300 python tools/bench_stage.py wgrad_tconv wgrad_proj wgrad_emb_c128; done > gpurun_out/${tag}_kb_fp32.log 2>&1; cat gpurun_out/${tag}_kb_fp32.log
