#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"conv_tc2_kernel" -c 1 -f -o gpurun_out/n1_full_conv_dproj_c64_fp32 python tools/bench_stage.py conv_dproj_c64 --once > gpurun_out/n1.log 2>&1; tail -1 gpurun_out/n1.log
