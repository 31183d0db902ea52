#!/bin/bash
tag=${1:-wg2}
mkdir -p gpurun_out
for d in 0 1 2 3; do echo "== AGCN_WG_DEBUG=$d fp32"; AGCN_WG_DEBUG=$d timeout 300 python tools/bench_stage.py wgrad_tconv wgrad_proj_c64; done > gpurun_out/${tag}_dbg_fp32.log 2>&1; cat gpurun_out/${tag}_dbg_fp32.log
for d in 0 2; do echo "== AGCN_WG_DEBUG=$d tf32"; AGCN_WG_DEBUG=$d timeout 300 python tools/bench_stage.py wgrad_tconv wgrad_proj_c64 --tf32; done > gpurun_out/${tag}_dbg_tf32.log 2>&1; cat gpurun_out/${tag}_dbg_tf32.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"wgrad_tc_kernel" -c 2 -f -o gpurun_out/${tag}_wg_tf32_full python tools/bench_stage.py wgrad_tconv_c128 wgrad_proj_c64 --tf32 --once > gpurun_out/${tag}_ncu1.log 2>&1; tail -1 gpurun_out/${tag}_ncu1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"wgrad_tc_kernel" -c 2 -f -o gpurun_out/${tag}_wg_fp32_full python tools/bench_stage.py wgrad_tconv_c128 wgrad_proj_c64 --once > gpurun_out/${tag}_ncu2.log 2>&1; tail -1 gpurun_out/${tag}_ncu2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel|conv_tc2_kernel" -c 5 -f -o gpurun_out/${tag}_conv_fp32_full python tools/bench_stage.py conv_emb_c64 conv_proj_c64 conv_dproj_c64 conv_tconv_c64 conv_proj_c256 --once > gpurun_out/${tag}_ncu3.log 2>&1; tail -1 gpurun_out/${tag}_ncu3.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_fp32.json 2>gpurun_out/${tag}_bench_fp32.err; tail -c 1300 gpurun_out/${tag}_bench_fp32.json
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --precision tf32 > gpurun_out/${tag}_bench_tf32.json 2>gpurun_out/${tag}_bench_tf32.err; tail -c 1300 gpurun_out/${tag}_bench_tf32.json
