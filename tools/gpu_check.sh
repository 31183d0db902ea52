#!/bin/bash
# One GPU-box pass: parity tests, a short bench in both precision modes, the ncu launch list of the bench command.
# usage: tools/gpu_check.sh <tag> [ncu-kernel-regex]
tag=${1:-run}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${tag}_pytest.log; echo "pytest rc=${PIPESTATUS[0]}" >> gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench_fp32.json 2> gpurun_out/${tag}_bench_fp32.err; tail -c 1500 gpurun_out/${tag}_bench_fp32.json
python bench.py --steps 5 --warmup 3 --precision tf32 --no-cpu-baseline > gpurun_out/${tag}_bench_tf32.json 2> gpurun_out/${tag}_bench_tf32.err; tail -c 600 gpurun_out/${tag}_bench_tf32.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
if [ -n "$2" ]; then
  ncu --set full --clock-control none --import-source on -k regex:$2 -s 30 -c 3 -f -o gpurun_out/${tag}_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_full.log 2>&1
  tail -3 gpurun_out/${tag}_ncu_full.log
fi
