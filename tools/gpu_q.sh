#!/bin/bash
tag=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${tag}_pytest.log; tail -3 gpurun_out/${tag}_pytest.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_fp32.json 2>gpurun_out/${tag}_bench_fp32.err; python tools/show_bench.py gpurun_out/${tag}_bench_fp32.json
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --precision tf32 > gpurun_out/${tag}_bench_tf32.json 2>gpurun_out/${tag}_bench_tf32.err; python tools/show_bench.py gpurun_out/${tag}_bench_tf32.json
for b in 32 16 8; do python bench.py --steps 5 --warmup 3 --no-cpu-baseline --precision tf32 --batch $b > gpurun_out/${tag}_bench_tf32_b$b.json 2>/dev/null; python tools/show_bench.py gpurun_out/${tag}_bench_tf32_b$b.json | head -1; done
