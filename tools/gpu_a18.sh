#!/bin/bash
tag=${1:-a18}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/${tag}_pytest.log; tail -4 gpurun_out/${tag}_pytest.log
python __graft_entry__.py --smoke 2>&1 | tail -1
( timeout 300 python tools/bench_stage.py wgrad gram mix_; echo "== tf32"; timeout 300 python tools/bench_stage.py wgrad gram mix_ --tf32 ) > gpurun_out/${tag}_stage.log 2>&1; cat gpurun_out/${tag}_stage.log
timeout 600 python bench.py --no-cpu-baseline --dump-kernels gpurun_out/${tag}_kernels_fp32.json > gpurun_out/${tag}_bench_fp32.json 2> gpurun_out/${tag}_bench_fp32.err; tail -3 gpurun_out/${tag}_bench_fp32.err
python tools/show_bench.py gpurun_out/${tag}_bench_fp32.json 2>/dev/null | head -1
python -c "
import json;d=json.loads(open('gpurun_out/${tag}_bench_fp32.json').read().strip().splitlines()[-1]);print('tf32_mode',d.get('tf32_mode',{}).get('value'), d.get('tf32_mode',{}).get('eager_value'));print('eager',d['eager_mode']['value'], 'graph_mode',d.get('graph_mode',{}).get('value'), d.get('graph_mode',{}).get('error'));print({k:v['share_of_step'] for k,v in d['entry_point_shares'].items()})"
