timeout 900 python -m pytest tests/test_gpu_stages.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
timeout 300 python tools/bench_stage.py gram_ mix_ conv_emb conv_proj conv_dproj conv_tconv 2>&1 | tee gpurun_out/r4a_stage.log
python bench.py --steps 10 --warmup 3 --dump-kernels gpurun_out/r4a_kernels.json > gpurun_out/r4a_bench.json 2> gpurun_out/r4a_bench.err; python -c "
import json;d=json.load(open('gpurun_out/r4a_bench.json'));print(d['value'],d['ms_per_step'],d['tf32_mode']['value'],d['e2e'])"
