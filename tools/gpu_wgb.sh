#!/bin/bash
tag=${1:-wgb}
mkdir -p gpurun_out
timeout 900 python tests/tools/tc_debug.py wg_1x1 wg_1x1_odd wg_taps9 wg_c256 wg_k768 wg_m384 wg_stride2 wg_res_stride2 wg_fc wg_big wg_big256 > gpurun_out/${tag}_tc_debug.log 2>&1; cat gpurun_out/${tag}_tc_debug.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${tag}_pytest.log; tail -3 gpurun_out/${tag}_pytest.log
timeout 600 python tools/bench_stage.py wgrad > gpurun_out/${tag}_stage_fp32.log 2>&1; cat gpurun_out/${tag}_stage_fp32.log
timeout 600 python tools/bench_stage.py wgrad --tf32 > gpurun_out/${tag}_stage_tf32.log 2>&1; cat gpurun_out/${tag}_stage_tf32.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_fp32.json 2>gpurun_out/${tag}_bench_fp32.err; python tools/show_bench.py gpurun_out/${tag}_bench_fp32.json 2>/dev/null | head -9; python -c "
import json;d=json.loads(open('gpurun_out/${tag}_bench_fp32.json').read().strip().splitlines()[-1]);print('tf32_mode',d.get('tf32_mode'))"
