#!/bin/bash
tag=${1:-r}
mkdir -p gpurun_out
timeout 900 python tests/tools/tc_debug.py k64 taps9 n256 stride2 dgrad_s2 wide_k big big256 wg_1x1 wg_taps9 wg_c256 wg_stride2 wg_big wg_big256 > gpurun_out/${tag}_tc_debug.log 2>&1; cat gpurun_out/${tag}_tc_debug.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${tag}_pytest.log; tail -3 gpurun_out/${tag}_pytest.log
timeout 600 python tools/bench_stage.py conv wgrad gram > gpurun_out/${tag}_stage_fp32.log 2>&1; cat gpurun_out/${tag}_stage_fp32.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_fp32.json 2>gpurun_out/${tag}_bench_fp32.err; python tools/show_bench.py gpurun_out/${tag}_bench_fp32.json
