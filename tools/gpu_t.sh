#!/bin/bash
# parity tests + smoke (+ optional extra command)
tag=${1:-t}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/${tag}_pytest.log; tail -12 gpurun_out/${tag}_pytest.log
python __graft_entry__.py --smoke 2>&1 | tail -2
shift
if [ -n "$1" ]; then eval "$@"; fi
