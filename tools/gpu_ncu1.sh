#!/bin/bash
tag=${1:-ncu1}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"wgrad_tc_kernel" -c 1 -f -o gpurun_out/${tag}_wg_fp32_full python tools/bench_stage.py wgrad_tconv_c128 --once > gpurun_out/${tag}_ncu.log 2>&1; tail -1 gpurun_out/${tag}_ncu.log
