#!/bin/bash
tag=${1:-multi}
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/${tag}_bench_fp32_2gpu.json 2> gpurun_out/${tag}_bench_fp32_2gpu.err; tail -2 gpurun_out/${tag}_bench_fp32_2gpu.err; python tools/show_bench.py gpurun_out/${tag}_bench_fp32_2gpu.json | head -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --impl reference > gpurun_out/${tag}_bench_ref_2gpu.json 2> gpurun_out/${tag}_bench_ref_2gpu.err; tail -c 600 gpurun_out/${tag}_bench_ref_2gpu.json
for w in utd mmact_imu utd_rgb; do timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload $w > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err; tail -2 gpurun_out/${tag}_bench_$w.err; python tools/show_bench.py gpurun_out/${tag}_bench_$w.json | head -3; done
