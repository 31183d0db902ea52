#!/bin/bash
mkdir -p gpurun_out
for m in fused plain; do
  if [ $m = plain ]; then export AGCN_NO_FUSED_STATS=1; fi
  python bench.py --precision tf32 --no-cpu-baseline --no-graph --dump-kernels gpurun_out/b1_kernels_tf32_$m.json > gpurun_out/b1_bench_tf32_$m.json 2> gpurun_out/b1_bench_tf32_$m.err
  python -c "
import json;d=json.loads(open('gpurun_out/b1_bench_tf32_$m.json').read().strip().splitlines()[-1]);print('$m', d['value'], d['ms_per_step']);print({k:v['share_of_step'] for k,v in d['entry_point_shares'].items()})"
done
