#!/bin/bash
# probe + per-kernel timings + one ncu full capture of a named bench_stage case
tag=${1:-probe}; case_name=${2:-gram_dg_c64}; kregex=${3:-joint_gram}
mkdir -p gpurun_out
timeout 120 tools/umma_rowoff_test > gpurun_out/${tag}_rowoff.log 2>&1; echo "rowoff rc=$?"
grep -c OK gpurun_out/${tag}_rowoff.log; grep -c MISMATCH gpurun_out/${tag}_rowoff.log
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/${tag}_pytest.log; tail -2 gpurun_out/${tag}_pytest.log
timeout 600 python tools/bench_stage.py > gpurun_out/${tag}_stage_fp32.log 2>&1; cat gpurun_out/${tag}_stage_fp32.log
timeout 600 python tools/bench_stage.py conv wgrad --tf32 > gpurun_out/${tag}_stage_tf32.log 2>&1; cat gpurun_out/${tag}_stage_tf32.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kregex -c 2 -f -o gpurun_out/${tag}_full python tools/bench_stage.py $case_name --once > gpurun_out/${tag}_ncu.log 2>&1; tail -2 gpurun_out/${tag}_ncu.log
