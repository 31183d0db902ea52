#!/bin/bash
tag=${1:-v3}
mkdir -p gpurun_out
timeout 900 python tests/tools/tc_debug.py > gpurun_out/${tag}_tc_debug.log 2>&1; cat gpurun_out/${tag}_tc_debug.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/${tag}_pytest.log; tail -3 gpurun_out/${tag}_pytest.log
timeout 600 python tools/bench_stage.py conv wgrad > gpurun_out/${tag}_stage_fp32.log 2>&1; cat gpurun_out/${tag}_stage_fp32.log
timeout 600 python tools/bench_stage.py conv wgrad --tf32 > gpurun_out/${tag}_stage_tf32.log 2>&1; cat gpurun_out/${tag}_stage_tf32.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_fp32.json 2>gpurun_out/${tag}_bench_fp32.err; tail -c 1300 gpurun_out/${tag}_bench_fp32.json
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --precision tf32 > gpurun_out/${tag}_bench_tf32.json 2>gpurun_out/${tag}_bench_tf32.err; tail -c 1300 gpurun_out/${tag}_bench_tf32.json
