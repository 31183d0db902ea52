#!/bin/bash
tag=${1:-gram}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_stages.py -m gpu -x -q -k "gram" 2>&1 | tail -15 > gpurun_out/${tag}_pytest_gram.log; cat gpurun_out/${tag}_pytest_gram.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${tag}_pytest.log; tail -5 gpurun_out/${tag}_pytest.log
timeout 300 python tools/bench_stage.py gram conv_dproj conv_emb > gpurun_out/${tag}_stage_fp32.log 2>&1; cat gpurun_out/${tag}_stage_fp32.log
timeout 300 python tools/bench_stage.py gram conv_dproj conv_emb conv_tconv_c256 --tf32 > gpurun_out/${tag}_stage_tf32.log 2>&1; cat gpurun_out/${tag}_stage_tf32.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_fp32.json 2>gpurun_out/${tag}_bench_fp32.err; tail -c 3000 gpurun_out/${tag}_bench_fp32.json; tail -3 gpurun_out/${tag}_bench_fp32.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --precision tf32 > gpurun_out/${tag}_bench_tf32.json 2>gpurun_out/${tag}_bench_tf32.err; tail -c 3000 gpurun_out/${tag}_bench_tf32.json
