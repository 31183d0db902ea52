#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stages.py -m gpu -q -x -k "conv or tf32_tensor" 2>&1 | tail -4
( echo "== fp32"; timeout 300 python tools/bench_stage.py conv_
  echo "== tf32"; timeout 300 python tools/bench_stage.py conv_ --tf32 ) > gpurun_out/a17_stage.log 2>&1; cat gpurun_out/a17_stage.log
