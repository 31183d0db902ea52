#!/bin/bash
tag=${1:-a6}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stages.py -m gpu -q -k "mix" 2>&1 | tail -5
( timeout 300 python tools/bench_stage.py mix_; echo "== tf32"; timeout 300 python tools/bench_stage.py mix_ --tf32 ) > gpurun_out/${tag}_stage.log 2>&1; cat gpurun_out/${tag}_stage.log
