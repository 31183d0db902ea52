"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel:  python tools/launch_summary.py file.csv [steps]   (steps omitted or 0: inferred from the once-per-step classifier-head kernel)"""
import collections
import csv
import re
import sys


def main(path, steps=1.0):
    rows = list(csv.reader(open(path)))
    hdr = None
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", d["Kernel Name"]).replace("void ", "")
        v = float(d["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}.get(d["Metric Unit"], 1e-6)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    # steps covered by the capture (-c N cuts the list mid-run): the classifier-head kernel launches once per step
    once = [v[0] for k, v in agg.items() if "linear_ce_fwd_kernel" in k]
    if steps <= 0:
        steps = float(once[0]) if once and once[0] > 0 else 1.0
    print(f"total {tot:.3f} ms over {sum(v[0] for v in agg.values())} launches; per step (/{steps:g}): {tot / steps:.3f} ms")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        print(f"{v[1] / steps:9.3f} ms/step {v[0] / steps:7.1f} launches/step {100 * v[1] / tot:5.1f}%  {k[:100]}")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.0)          # 0: infer the step count from the launch list
