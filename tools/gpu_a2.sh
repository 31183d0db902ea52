#!/bin/bash
# round-2 experiment pass: parity tests, the touched stages in isolation (new vs old path), bench with graph mode
tag=${1:-a2}
mkdir -p gpurun_out
t0=$SECONDS
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${tag}_pytest.log; tail -4 gpurun_out/${tag}_pytest.log; echo "pytest $((SECONDS-t0))s"
( echo "== new"; timeout 300 python tools/bench_stage.py first_unit res_dgrad mix_score_bwd
  echo "== old (AGCN_SKINNY_SCALAR AGCN_TC2_NO_SKIP_PARITY AGCN_MIX_SCORE_SIMT)"; AGCN_SKINNY_SCALAR=1 AGCN_TC2_NO_SKIP_PARITY=1 AGCN_MIX_SCORE_SIMT=1 timeout 300 python tools/bench_stage.py first_unit_wgrad res_dgrad mix_score_bwd
  echo "== new tf32"; timeout 300 python tools/bench_stage.py res_dgrad mix_score_bwd --tf32 ) > gpurun_out/${tag}_stage.log 2>&1; cat gpurun_out/${tag}_stage.log
timeout 600 python bench.py --dump-kernels gpurun_out/${tag}_kernels_fp32.json > gpurun_out/${tag}_bench_fp32.json 2> gpurun_out/${tag}_bench_fp32.err; tail -3 gpurun_out/${tag}_bench_fp32.err
python tools/show_bench.py gpurun_out/${tag}_bench_fp32.json 2>/dev/null | head -4
python -c "
import json;d=json.loads(open('gpurun_out/${tag}_bench_fp32.json').read().strip().splitlines()[-1]);print('tf32_mode',d.get('tf32_mode',{}).get('value'));print('graph_mode',d.get('graph_mode'))"
timeout 300 python bench.py --mode infer --steps 3 > gpurun_out/${tag}_bench_infer.json 2> gpurun_out/${tag}_bench_infer.err; tail -3 gpurun_out/${tag}_bench_infer.err; head -c 900 gpurun_out/${tag}_bench_infer.json
