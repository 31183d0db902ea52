#!/bin/bash
tag=${1:-epi}
mkdir -p gpurun_out
timeout 900 python tests/tools/tc_debug.py k32 k64 multi_tile taps9 n256 n96_k16 v20_acc v22 stride2 res_stride2 dgrad_s1 dgrad_s2 wide_k big big256 > gpurun_out/${tag}_tc_debug.log 2>&1; cat gpurun_out/${tag}_tc_debug.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/${tag}_pytest.log; tail -3 gpurun_out/${tag}_pytest.log
timeout 600 python tools/bench_stage.py conv wgrad_tconv_c64 > gpurun_out/${tag}_stage_fp32.log 2>&1; cat gpurun_out/${tag}_stage_fp32.log
timeout 600 python tools/bench_stage.py conv --tf32 > gpurun_out/${tag}_stage_tf32.log 2>&1; cat gpurun_out/${tag}_stage_tf32.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_fp32.json 2>gpurun_out/${tag}_bench_fp32.err; tail -c 1300 gpurun_out/${tag}_bench_fp32.json
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --precision tf32 > gpurun_out/${tag}_bench_tf32.json 2>gpurun_out/${tag}_bench_tf32.err; tail -c 1300 gpurun_out/${tag}_bench_tf32.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
