#!/bin/bash
# end-of-round measurement pass: parity tests, smoke, bench lines (all workloads / modes), stage timings, ncu launch list,
# ncu --set full of the dominant kernels.  usage: tools/gpu_final.sh <tag>
tag=${1:-r2z}
mkdir -p gpurun_out
o=gpurun_out/${tag}
t0=$SECONDS
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > ${o}_pytest.log; tail -2 ${o}_pytest.log
python __graft_entry__.py --smoke 2>&1 | tail -1 | tee ${o}_smoke.log
echo "tests $((SECONDS-t0))s"; t0=$SECONDS
timeout 600 python bench.py --dump-kernels ${o}_kernels_fp32.json > ${o}_bench_fp32.json 2> ${o}_bench_fp32.err; tail -2 ${o}_bench_fp32.err
timeout 600 python bench.py --precision tf32 --no-cpu-baseline --dump-kernels ${o}_kernels_tf32.json > ${o}_bench_tf32.json 2> ${o}_bench_tf32.err; tail -2 ${o}_bench_tf32.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > ${o}_bench_ref.json 2> ${o}_bench_ref.err
for w in utd mmact_imu utd_rgb; do timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload $w > ${o}_bench_$w.json 2> ${o}_bench_$w.err; tail -2 ${o}_bench_$w.err; done
timeout 600 python bench.py --mode infer --steps 3 > ${o}_bench_infer_n256.json 2> ${o}_bench_infer_n256.err; tail -2 ${o}_bench_infer_n256.err
timeout 600 python bench.py --mode infer --steps 2 --batch 1024 --precision tf32 > ${o}_bench_infer_n1024_tf32.json 2> ${o}_bench_infer_n1024_tf32.err; tail -2 ${o}_bench_infer_n1024_tf32.err
for f in fp32 tf32 utd mmact_imu utd_rgb; do python tools/show_bench.py ${o}_bench_$f.json 2>/dev/null | head -1; done
python -c "
import json
for f in ('infer_n256','infer_n1024_tf32','ref'):
    d=json.loads(open('${o}_bench_'+f+'.json').read().strip().splitlines()[-1]); print(f, d['value'], d.get('ms_per_step'), (d.get('model_roofline') or {}).get('achieved_frac_of_hbm_ceiling'))"
echo "bench $((SECONDS-t0))s"; t0=$SECONDS
timeout 400 python tools/bench_stage.py > ${o}_stage_fp32.log 2>&1; timeout 400 python tools/bench_stage.py --tf32 > ${o}_stage_tf32.log 2>&1
echo "stage $((SECONDS-t0))s"; t0=$SECONDS
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file ${o}_launches_fp32.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-tf32 --no-graph > ${o}_ncu_bench.log 2>&1; echo "ncu list $((SECONDS-t0))s"; t0=$SECONDS
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"wgrad_tc_kernel" -c 1 -f -o ${o}_full_wgrad_c256_fp32 python tools/bench_stage.py wgrad_tconv_c256 --once > ${o}_ncu_wg.log 2>&1; tail -1 ${o}_ncu_wg.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"conv_tc2_kernel" -c 1 -f -o ${o}_full_conv_c256_fp32 python tools/bench_stage.py conv_tconv_c256 --once > ${o}_ncu_cv.log 2>&1; tail -1 ${o}_ncu_cv.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"conv_tc2_kernel" -c 1 -f -o ${o}_full_conv_c64_fp32 python tools/bench_stage.py conv_tconv_c64 --once > ${o}_ncu_cv64.log 2>&1; tail -1 ${o}_ncu_cv64.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"wgrad_tc_kernel" -c 1 -f -o ${o}_full_wgrad_c64_fp32 python tools/bench_stage.py wgrad_tconv_c64 --once > ${o}_ncu_wg64.log 2>&1; tail -1 ${o}_ncu_wg64.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"conv_tc2_kernel" -c 1 -f -o ${o}_full_conv_dproj_c64_tf32 python tools/bench_stage.py conv_dproj_c64 --once --tf32 > ${o}_ncu_dp.log 2>&1; tail -1 ${o}_ncu_dp.log
echo "ncu full $((SECONDS-t0))s"
