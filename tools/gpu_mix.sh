#!/bin/bash
tag=${1:-mix}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_stages.py -m gpu -x -q -k "mix" 2>&1 | tail -15 > gpurun_out/${tag}_pytest_mix.log; tail -12 gpurun_out/${tag}_pytest_mix.log | cut -c1-300
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${tag}_pytest.log; tail -3 gpurun_out/${tag}_pytest.log
timeout 300 python tools/bench_stage.py mix_fwd mix_bwd > gpurun_out/${tag}_stage_fp32.log 2>&1; cat gpurun_out/${tag}_stage_fp32.log
timeout 300 python tools/bench_stage.py mix_fwd mix_bwd --tf32 > gpurun_out/${tag}_stage_tf32.log 2>&1; cat gpurun_out/${tag}_stage_tf32.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_fp32.json 2>gpurun_out/${tag}_bench_fp32.err; python tools/show_bench.py gpurun_out/${tag}_bench_fp32.json
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --precision tf32 > gpurun_out/${tag}_bench_tf32.json 2>gpurun_out/${tag}_bench_tf32.err; python tools/show_bench.py gpurun_out/${tag}_bench_tf32.json
