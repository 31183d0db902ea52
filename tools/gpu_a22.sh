#!/bin/bash
mkdir -p gpurun_out
export AGCN_SPLIT_IMPLICIT_HI=1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12
python __graft_entry__.py --smoke 2>&1 | tail -1
( timeout 300 python tools/bench_stage.py conv wgrad gram mix_ ) > gpurun_out/a22_stage.log 2>&1; cat gpurun_out/a22_stage.log
