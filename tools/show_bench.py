"""Pretty-print a bench.py JSON line:  python tools/show_bench.py file.json"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(f"{d.get('config', {}).get('precision_mode')} value {d['value']} seq/s  {d['ms_per_step']} ms/step  e2e {d['e2e']['value']}  launches {d.get('gpu_launches')}  "
      f"hbm-ceiling frac {d.get('model_roofline', {}).get('achieved_frac_of_hbm_ceiling')}")
r = d.get("roofline") or {}
print("  roofline:", {k: r.get(k) for k in ("bound", "achieved", "peak", "frac", "traffic", "kernel", "share_of_step")})
for t in d.get("top_kernels", []):
    print(f"   {t['share_of_step']:.4f} {t['avg_launch_ms']:8.4f} ms x{t['launches_timed']:3d} {t['tflops']:8.2f} TF/s {t['gbs']:8.1f} GB/s  {t['kernel']}")
