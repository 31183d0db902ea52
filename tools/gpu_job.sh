#!/bin/bash
# One GPU-box job: tools/gpu_job.sh <tag> [steps...]   steps: pytest | pytestnew | peaks | bench | stage | stagetf32
tag=$1; shift
mkdir -p gpurun_out
for step in "$@"; do
  case $step in
    pytest)    python -m pytest tests -m gpu -q -rf -p no:cacheprovider > gpurun_out/${tag}_pytest.log 2>&1; tail -5 gpurun_out/${tag}_pytest.log ;;
    pytestnew) python -m pytest tests/test_gpu_baseline_shapes.py -m gpu -q -rf -s -p no:cacheprovider > gpurun_out/${tag}_pytest_new.log 2>&1; tail -5 gpurun_out/${tag}_pytest_new.log ;;
    peaks)     python tools/measure_peaks.py > gpurun_out/${tag}_peaks.json 2> gpurun_out/${tag}_peaks.err; cat gpurun_out/${tag}_peaks.json ;;
    bench)     python bench.py --steps 10 --warmup 3 --dump-kernels gpurun_out/${tag}_kernels.json > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 600 gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err ;;
    benchbf)   python bench.py --steps 10 --warmup 3 --precision bf16x3 --no-cpu-baseline --dump-kernels gpurun_out/${tag}_kernels_bf16x3.json > gpurun_out/${tag}_bench_bf16x3.json 2> gpurun_out/${tag}_bench_bf16x3.err; tail -c 400 gpurun_out/${tag}_bench_bf16x3.json; tail -3 gpurun_out/${tag}_bench_bf16x3.err ;;
    stage)     python tools/bench_stage.py > gpurun_out/${tag}_stage_fp32.log 2>&1; tail -3 gpurun_out/${tag}_stage_fp32.log ;;
    stagebf)   python tools/bench_stage.py conv wgrad --bf16x3 > gpurun_out/${tag}_stage_bf16x3.log 2>&1; cat gpurun_out/${tag}_stage_bf16x3.log ;;
    stageconv) python tools/bench_stage.py conv wgrad > gpurun_out/${tag}_stage_fp32.log 2>&1; cat gpurun_out/${tag}_stage_fp32.log ;;
    convtest)  timeout 600 python -m pytest tests/test_gpu_stages.py -m gpu -q -rf -x -k "conv" -p no:cacheprovider > gpurun_out/${tag}_convtest.log 2>&1; tail -15 gpurun_out/${tag}_convtest.log ;;
    stagetf32) python tools/bench_stage.py --tf32 > gpurun_out/${tag}_stage_tf32.log 2>&1 ;;
    smoke)     python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -2 gpurun_out/${tag}_smoke.log ;;
    ncu)       # NCU_CASES="case[:flag] ..." NCU_KERNEL=regex : one --set full capture per case, raw page as CSV
               for spec in $NCU_CASES; do
                 c=${spec%%:*}; f=""; [ "$spec" != "$c" ] && f="--${spec#*:}"
                 out=gpurun_out/${tag}_ncu_${c}${f#--}
                 timeout 300 ncu --set full --clock-control none --import-source on -k regex:${NCU_KERNEL:-tc} -c ${NCU_COUNT:-1} -f -o $out python tools/bench_stage.py $c --once $f > ${out}.log 2>&1
                 ncu -i ${out}.ncu-rep --page raw --csv > ${out}_raw.csv 2>/dev/null; rm -f ${out}.ncu-rep
                 tail -2 ${out}.log
               done ;;
    optim)     python -m pytest tests/test_gpu_optim.py -m gpu -q -rf -x -p no:cacheprovider > gpurun_out/${tag}_optim.log 2>&1; tail -8 gpurun_out/${tag}_optim.log ;;
    pytests)   python -m pytest tests -m gpu -q -rf -s -p no:cacheprovider > gpurun_out/${tag}_pytest_s.log 2>&1; grep -E "^\[|passed|failed" gpurun_out/${tag}_pytest_s.log | tail -60 ;;
    launcher)  python -m pytest tests/test_gpu_launcher.py -m gpu -q -rf -s -x -p no:cacheprovider > gpurun_out/${tag}_launcher.log 2>&1; tail -12 gpurun_out/${tag}_launcher.log ;;
    benchN)    # BENCH_N=<gpus> [BENCH_ARGS=...]: the driver's multi-GPU launch
               python -m torch.distributed.run --nnodes=1 --nproc-per-node ${BENCH_N:-2} --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus ${BENCH_N:-2} --steps 10 --warmup 3 ${BENCH_ARGS} > gpurun_out/${tag}_bench_n${BENCH_N:-2}.json 2> gpurun_out/${tag}_bench_n${BENCH_N:-2}.err; tail -c 1500 gpurun_out/${tag}_bench_n${BENCH_N:-2}.json; tail -5 gpurun_out/${tag}_bench_n${BENCH_N:-2}.err ;;
    graphconv) python -m pytest tests/test_graphconv.py -m gpu -q -rf -x -p no:cacheprovider > gpurun_out/${tag}_graphconv.log 2>&1; tail -12 gpurun_out/${tag}_graphconv.log ;;
    probes)    # PROBE_CASES="case ..." PROBE_DBG="0 1 8 ..." : AGCN_CONV_DEBUG limiter probes on the -DAGCN_PROBES build (timing only)
               for d in $PROBE_DBG; do echo "== AGCN_CONV_DEBUG=$d"; AGCN_B200_LIB=$PWD/fusion_gcn_b200/libagcn_b200_probes.so AGCN_CONV_DEBUG=$d python tools/bench_stage.py $PROBE_CASES ${PROBE_FLAGS}; done > gpurun_out/${tag}_probes.log 2>&1; cat gpurun_out/${tag}_probes.log ;;
    infer)     for n in 256 1024; do for pr in fp32 tf32; do python bench.py --mode infer --batch $n --precision $pr --steps 5 --warmup 3 > gpurun_out/${tag}_bench_infer_n${n}_${pr}.json 2> gpurun_out/${tag}_bench_infer.err; python -c "import json;d=json.load(open('gpurun_out/${tag}_bench_infer_n${n}_${pr}.json'));print('infer',$n,'$pr',d['value'],d['model_roofline']['achieved_frac_of_hbm_ceiling'])"; done; done ;;
    workloads) for w in utd mmact_imu utd_rgb; do python bench.py --workload $w --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_${w}.json 2> gpurun_out/${tag}_bench_${w}.err; python -c "import json;d=json.load(open('gpurun_out/${tag}_bench_${w}.json'));print('$w',d['value'],d['tf32_mode']['value'])"; done ;;
    launches)  # ncu launch list of one short bench run (a number printed under ncu is never a bench value) + per-kernel summary
               ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-tf32 --no-graph > gpurun_out/${tag}_launches_bench.log 2>&1
               python tools/launch_summary.py gpurun_out/${tag}_launches.csv 0 > gpurun_out/${tag}_launch_summary.txt 2>&1; head -14 gpurun_out/${tag}_launch_summary.txt ;;
    *) echo "unknown step $step" ;;
  esac
done
