#!/bin/bash
tag=${1:-m2}
mkdir -p gpurun_out
nvidia-smi -L | head -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/${tag}_bench_fp32_2gpu.json 2> gpurun_out/${tag}_bench_fp32_2gpu.err; tail -3 gpurun_out/${tag}_bench_fp32_2gpu.err
python -c "
import json;d=json.loads(open('gpurun_out/${tag}_bench_fp32_2gpu.json').read().strip().splitlines()[-1]);print(d['value'], d['ms_per_step'], d['e2e'], d['config']['launch_mode']);print('eager',d['eager_mode']['value'],'graph',d.get('graph_mode'));print('tf32',d['tf32_mode']['value'])"
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_fp32_1gpu.json 2> gpurun_out/${tag}_bench_fp32_1gpu.err; tail -2 gpurun_out/${tag}_bench_fp32_1gpu.err
python -c "
import json;d=json.loads(open('gpurun_out/${tag}_bench_fp32_1gpu.json').read().strip().splitlines()[-1]);print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['launch_mode'], d['gpu_launches']);print('eager',d['eager_mode']['value'])"
