#!/bin/bash
# wgrad bring-up: correctness cases (each in its own subprocess under a timeout), stage timings, eager-PyTorch baseline, ncu of the joint kernels
tag=${1:-wg}
mkdir -p gpurun_out
timeout 900 python tests/tools/tc_debug.py wg_1x1 wg_1x1_odd wg_taps9 wg_c256 wg_k768 wg_m384 wg_stride2 wg_res_stride2 wg_fc wg_big wg_big256 > gpurun_out/${tag}_tc_debug.log 2>&1; cat gpurun_out/${tag}_tc_debug.log
timeout 600 python -m pytest tests/test_gpu_stages.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/${tag}_pytest_stages.log; tail -3 gpurun_out/${tag}_pytest_stages.log
timeout 600 python tools/bench_stage.py wgrad > gpurun_out/${tag}_stage_fp32.log 2>&1; cat gpurun_out/${tag}_stage_fp32.log
timeout 600 python tools/bench_stage.py wgrad --tf32 > gpurun_out/${tag}_stage_tf32.log 2>&1; cat gpurun_out/${tag}_stage_tf32.log
timeout 600 python tests/tools/bench_ref_gpu.py > gpurun_out/${tag}_ref_gpu.log 2>&1; timeout 600 python tests/tools/bench_ref_gpu.py --tf32 >> gpurun_out/${tag}_ref_gpu.log 2>&1; cat gpurun_out/${tag}_ref_gpu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"joint_gram|joint_mix" -c 6 -f -o gpurun_out/${tag}_joint_full python tools/bench_stage.py gram_dg_c64 mix_fwd_c64 mix_bwd_c64 mix_score_bwd_c64 gram_score_c64 --once > gpurun_out/${tag}_ncu_joint.log 2>&1; tail -2 gpurun_out/${tag}_ncu_joint.log
