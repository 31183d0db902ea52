#!/bin/bash
mkdir -p gpurun_out
o=gpurun_out/r2z
( echo "== AGCN_TC2_NLO=1 fp32"; AGCN_TC2_NLO=1 timeout 200 python tools/bench_stage.py conv_emb_c64 conv_proj_c64 conv_dproj_c64 conv_dproj_c128
  echo "== default"; timeout 200 python tools/bench_stage.py conv_emb_c64 conv_proj_c64 conv_dproj_c64 conv_dproj_c128 ) > gpurun_out/n2_nlo1.log 2>&1; cat gpurun_out/n2_nlo1.log
timeout 200 ncu --set full --clock-control none -k regex:"gram_tc_kernel" -c 1 -f -o ${o}_full_gram_dg_c64_tf32 python tools/bench_stage.py gram_dg_c64 --once --tf32 > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none -k regex:"mix_tc_kernel" -c 1 -f -o ${o}_full_mix_fwd_c64_tf32 python tools/bench_stage.py mix_fwd_c64 --once --tf32 > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none -k regex:"wgrad_tc_kernel" -c 1 -f -o ${o}_full_wgrad_proj_c64_tf32 python tools/bench_stage.py wgrad_proj_c64 --once --tf32 > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none -k regex:"conv_tc2_kernel" -c 1 -f -o ${o}_full_conv_tconv_c128_tf32 python tools/bench_stage.py conv_tconv_c128 --once --tf32 > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none -k regex:"bn_bwd_apply_kernel" -c 1 -f -o ${o}_full_bn_bwd_apply_c64 python tools/bench_stage.py bn_c64 --once > /dev/null 2>&1
ls gpurun_out/*.ncu-rep | tail -6
