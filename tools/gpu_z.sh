#!/bin/bash
tag=${1:-r2z}
mkdir -p gpurun_out
o=gpurun_out/${tag}
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > ${o}_pytest.log; tail -2 ${o}_pytest.log
python __graft_entry__.py --smoke 2>&1 | tail -1 | tee ${o}_smoke.log
timeout 600 python bench.py --dump-kernels ${o}_kernels_fp32.json > ${o}_bench_fp32.json 2> ${o}_bench_fp32.err; tail -2 ${o}_bench_fp32.err
timeout 600 python bench.py --precision tf32 --no-cpu-baseline --dump-kernels ${o}_kernels_tf32.json > ${o}_bench_tf32.json 2> ${o}_bench_tf32.err; tail -2 ${o}_bench_tf32.err
for f in fp32 tf32; do python tools/show_bench.py ${o}_bench_$f.json 2>/dev/null | head -2; done
timeout 400 python tools/bench_stage.py > ${o}_stage_fp32.log 2>&1; timeout 400 python tools/bench_stage.py --tf32 > ${o}_stage_tf32.log 2>&1
