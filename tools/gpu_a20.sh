#!/bin/bash
mkdir -p gpurun_out
( for n in 2 3 4; do echo "== AGCN_TC2_NLO=$n fp32"; AGCN_TC2_NLO=$n timeout 300 python tools/bench_stage.py conv_emb conv_proj conv_dproj; done ) > gpurun_out/a20_stage.log 2>&1; cat gpurun_out/a20_stage.log
timeout 600 python -m pytest tests/test_gpu_stages.py -m gpu -q -x -k "conv" 2>&1 | tail -2
