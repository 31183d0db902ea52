#!/bin/bash
# validation + measurement pass: parity tests, bench (full JSON + per-signature dump), ncu launch list, ncu --set full of the two dominant kernels
tag=${1:-a1}
mkdir -p gpurun_out
t0=$SECONDS
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/${tag}_pytest.log; tail -3 gpurun_out/${tag}_pytest.log; echo "pytest $((SECONDS-t0))s"
t0=$SECONDS
timeout 600 python bench.py --dump-kernels gpurun_out/${tag}_kernels_fp32.json > gpurun_out/${tag}_bench_fp32.json 2> gpurun_out/${tag}_bench_fp32.err; tail -3 gpurun_out/${tag}_bench_fp32.err
python tools/show_bench.py gpurun_out/${tag}_bench_fp32.json 2>/dev/null | head -12; echo "bench $((SECONDS-t0))s"
t0=$SECONDS
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${tag}_launches_fp32.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-tf32 > gpurun_out/${tag}_ncu_bench.log 2>&1; echo "ncu list $((SECONDS-t0))s"
t0=$SECONDS
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"wgrad_tc_kernel" -c 1 -f -o gpurun_out/${tag}_full_wgrad_c256_fp32 python tools/bench_stage.py wgrad_tconv_c256 --once > gpurun_out/${tag}_ncu_wg.log 2>&1; tail -1 gpurun_out/${tag}_ncu_wg.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"conv_tc2_kernel" -c 1 -f -o gpurun_out/${tag}_full_conv_c256_fp32 python tools/bench_stage.py conv_tconv_c256 --once > gpurun_out/${tag}_ncu_cv.log 2>&1; tail -1 gpurun_out/${tag}_ncu_cv.log
echo "ncu full $((SECONDS-t0))s"
