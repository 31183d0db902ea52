"""Per-kernel DRAM traffic and time from `ncu --set full` raw pages -> profiles/ncu_traffic.json (what bench.py reports as roofline.traffic).

    python tools/ncu_traffic.py profiles/<tag>_ncu_<case>_raw.csv ...  [--summary profiles/<tag>_ncu_summary.md]

The raw CSVs come from `tools/gpu_job.sh <tag> ncu` (one `--set full` capture of the first tensor-core kernel of a
tools/bench_stage.py case); the case name in the file name selects the bench.py launch signature(s) the capture stands for.
"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NB, V = 128, 25
SHAPES = {"c64": (64, 300), "c128": (128, 150), "c256": (256, 75)}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3,
        "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}
KEEP = ["sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct"]


def signatures(case):
    """bench.py launch signatures (ops.py `sig=`) a bench_stage.py case corresponds to."""
    m = re.match(r"(conv|wgrad|wgradpre)_(emb|proj|dproj|tconv)_(c\d+)$", case)
    if m:
        kind, nm, tag = m.groups()
        if kind == "wgradpre":          # the pre-split form is what the step runs for the temporal convolutions (alias agcn_conv_wgrad)
            kind = "wgrad"
        c, t = SHAPES[tag]
        cin, cout, taps = {"emb": (c, 6 * (c // 4), 1), "proj": (3 * c, c, 1), "dproj": (c, 3 * c, 1), "tconv": (c, c, 9)}[nm]
        if kind == "wgrad":
            return [f"agcn_conv_wgrad[{NB}, {t}, {t}, {V}, {cin}, {cout}, {taps}, 1]"]
        # forward gather; the input-gradient launch of the same shape (transposed = 1) moves the same bytes
        return [f"agcn_conv_fwd[{NB}, {t}, {t}, {V}, {cin}, {cout}, {taps}, 1, {tr}, 0]" for tr in ((0, 1) if taps == 9 else (0,))]
    m = re.match(r"gram_dg_(c\d+)$", case)
    if m:
        c, t = SHAPES[m.group(1)]
        return [f"agcn_joint_gram[{NB}, {t}, {V}, {c}, {3 * c}, 3, {c}, 1]"]
    m = re.match(r"mix_fwd_(c\d+)$", case)
    if m:
        c, t = SHAPES[m.group(1)]
        return [f"agcn_joint_mix[{NB}, {t}, {V}, {c}, {3 * c}, {c}, 0, 0]"]
    return []


def parse(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    out = []
    for r in rows[hdr + 2:]:
        if len(r) != len(names):
            continue
        d = {}
        for n, u, v in zip(names, units, r):
            try:
                d[n] = float(v.replace(",", "")) * UNIT.get(u, 1.0) if n.startswith(("dram__bytes", "gpu__time_duration")) else v
            except ValueError:
                d[n] = v
        out.append(d)
    return out


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    summary = sys.argv[sys.argv.index("--summary") + 1] if "--summary" in sys.argv else None
    args = [a for a in args if a != summary]
    db_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    db = json.load(open(db_path)) if os.path.isfile(db_path) else {}
    lines = ["| case | kernel | time (ms, under ncu) | DRAM read + write (MB) | SM % | DRAM % | L2 % | tensor pipe % | shared-memory LSU % | regs |", "|---|---|---|---|---|---|---|---|---|---|"]
    for path in args:
        case = re.sub(r".*_ncu_", "", os.path.basename(path)).replace("_raw.csv", "")
        flag = ""
        for f in ("tf32", "bf16x3"):
            if case.endswith(f):
                case, flag = case[:-len(f)], f
        for k in parse(path):
            dram = k.get("dram__bytes_read.sum", 0.0) + k.get("dram__bytes_write.sum", 0.0)
            ms = k.get("gpu__time_duration.sum", float("nan"))
            kern = re.sub(r"\(.*", "", str(k.get("Kernel Name", "?"))).replace("void ", "")
            lines.append(f"| {case}{' (' + flag + ')' if flag else ''} | `{kern}` | {ms:.4f} | {dram / 1e6:.1f} | " + " | ".join(
                str(k.get(m, "")) for m in KEEP[:5]) + f" | {k.get(KEEP[5], '')} |")
            if not flag:                                   # bench.py's headline is the strict fp32 mode
                for sig in signatures(case):
                    db[sig] = {"dram_bytes": int(dram), "ncu_time_ms": round(ms, 5),
                               "source": f"{os.path.relpath(path, ROOT)} (ncu --set full, {kern}, dram__bytes_read.sum + dram__bytes_write.sum)"}
    json.dump(db, open(db_path, "w"), indent=1)
    text = "\n".join(lines)
    print(text)
    if summary:
        open(summary, "w").write("# ncu --set full summaries (one capture per case, cold cache, serialised -- shares and bytes, not bench times)\n\n" + text + "\n")


if __name__ == "__main__":
    main()
