#!/bin/bash
tag=${1:-m4}; n=${2:-4}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/${tag}_bench_fp32_${n}gpu.json 2> gpurun_out/${tag}_bench_fp32_${n}gpu.err; tail -3 gpurun_out/${tag}_bench_fp32_${n}gpu.err
python -c "
import json;d=json.loads(open('gpurun_out/${tag}_bench_fp32_${n}gpu.json').read().strip().splitlines()[-1]);print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['launch_mode']);print('eager',d['eager_mode']['value'],'tf32',d['tf32_mode']['value'], d['tf32_mode']['launch_mode'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $n --steps 2 --warmup 1 --impl reference > gpurun_out/${tag}_bench_ref_${n}gpu.json 2> gpurun_out/${tag}_bench_ref_${n}gpu.err; tail -c 300 gpurun_out/${tag}_bench_ref_${n}gpu.json
