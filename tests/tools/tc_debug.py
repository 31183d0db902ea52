"""Bring-up harness for the tcgen05 implicit-GEMM path (run on the GPU box):  python tests/tools/tc_debug.py [case ...]
Each case runs in its own subprocess under a timeout so that a deadlocked kernel cannot take the whole call down.
Compares agcn_conv_fwd(precision=TF32) with fp64 contractions of (a) the raw fp32 inputs, (b) inputs truncated to
TF32, (c) inputs rounded to TF32, and prints where the largest errors sit."""
import subprocess
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

CASES = {  # name: nb, t_in, v, cin, cout, taps, stride, transposed, bias, accumulate
    "k32": (1, 5, 25, 32, 64, 1, 1, 0, 0, 0),
    "k64": (1, 5, 25, 64, 64, 1, 1, 0, 1, 0),
    "multi_tile": (3, 23, 25, 64, 64, 1, 1, 0, 1, 0),
    "taps9": (2, 20, 25, 64, 64, 9, 1, 0, 1, 0),
    "n256": (2, 20, 25, 128, 256, 9, 1, 0, 1, 0),
    "n96_k16": (2, 12, 25, 16, 96, 1, 1, 0, 1, 0),
    "v20_acc": (2, 14, 20, 64, 128, 9, 1, 0, 1, 1),
    "v22": (2, 14, 22, 32, 32, 9, 1, 0, 0, 0),
    "stride2": (2, 21, 25, 64, 128, 9, 2, 0, 1, 0),
    "res_stride2": (2, 20, 25, 64, 128, 1, 2, 0, 1, 0),
    "dgrad_s1": (2, 20, 25, 64, 64, 9, 1, 1, 0, 0),
    "dgrad_s2": (2, 21, 25, 128, 64, 9, 2, 1, 0, 1),
    "wide_k": (1, 10, 25, 768, 256, 1, 1, 0, 0, 0),
    "big": (128, 300, 25, 64, 64, 9, 1, 0, 1, 0),
    "big256": (128, 75, 25, 256, 256, 9, 1, 0, 1, 0),
}


WG_CASES = {  # name: nb, t_in, v, cin, cout, taps, stride
    "wg_1x1": (2, 12, 25, 64, 64, 1, 1),
    "wg_1x1_odd": (3, 13, 20, 48, 96, 1, 1),
    "wg_taps9": (2, 20, 25, 64, 64, 9, 1),
    "wg_c256": (2, 20, 25, 256, 256, 9, 1),
    "wg_k768": (2, 10, 25, 768, 256, 1, 1),
    "wg_m384": (2, 10, 25, 256, 384, 1, 1),
    "wg_stride2": (2, 21, 25, 64, 128, 9, 2),
    "wg_res_stride2": (2, 20, 22, 64, 128, 1, 2),
    "wg_fc": (1, 1, 64, 256, 60, 1, 1),
    "wg_big": (128, 300, 25, 64, 64, 9, 1),
    "wg_big256": (128, 75, 25, 256, 256, 9, 1),
}


def run_wg_case(name):
    import torch
    from fusion_gcn_b200 import ops as K
    from oracle import stages as S
    nb, t_in, v, cin, cout, taps, stride = WG_CASES[name]
    pad = (taps - 1) // 2
    t_out = (t_in + 2 * pad - taps) // stride + 1
    g = torch.Generator().manual_seed(0)
    x = torch.randn(nb, t_in, v, cin, generator=g)
    dy = torch.randn(nb, t_out, v, cout, generator=g)
    kw = dict(taps=taps, stride=stride, pad=pad)
    xc, dyc = x.cuda(), dy.cuda()
    dw, db = K.conv_wgrad(dyc, xc, precision=K.PREC_TF32, **kw)
    torch.cuda.synchronize()
    dw3, db3 = K.conv_wgrad(dyc, xc, precision=K.PREC_FP32, **kw)
    torch.cuda.synchronize()
    big = nb * t_out * v > 100000

    def trunc(t_):
        return (t_.view(torch.int32) & ~0x1FFF).view(torch.float32)

    if big:
        ref_raw, db_ref = K.conv_wgrad(dyc, xc, precision=K.PREC_FP32_FFMA, **kw)
        ref_raw, db_ref = ref_raw.double().cpu(), db_ref.double().cpu()
        ref_tr = K.conv_wgrad(trunc(dy).cuda(), trunc(x).cuda(), precision=K.PREC_FP32_FFMA, **kw)[0].double().cpu()
    else:
        ref_raw, db_ref = S.conv_wgrad(dy.double(), x.double(), **kw)
        ref_tr = S.conv_wgrad(trunc(dy).double(), trunc(x).double(), **kw)[0]
    d = dw.double().cpu()
    e_raw = ((d - ref_raw).abs().max() / ref_raw.abs().max()).item()
    e_tr = ((d - ref_tr).abs().max() / ref_tr.abs().max()).item()
    e_b = ((db.double().cpu() - db_ref).abs().max() / db_ref.abs().max()).item()
    e3 = ((dw3.double().cpu() - ref_raw).abs().max() / ref_raw.abs().max()).item()
    print(f"[{name}] tf32 dw: vs raw {e_raw:.3e} | vs truncated-TF32 {e_tr:.3e} | dbias {e_b:.3e} || 3xTF32 dw vs raw {e3:.3e}")
    if e_raw > 5e-3:
        err = (d - ref_raw).abs() > 1e-2 * ref_raw.abs().max()
        print("   bad co:", err.any(dim=2).any(dim=1).nonzero().flatten()[:32].tolist())
        print("   bad tap:", err.any(dim=2).any(dim=0).nonzero().flatten().tolist())
        print("   bad ci:", err.any(dim=1).any(dim=0).nonzero().flatten()[:32].tolist())
        print("   dw[0,0,:8] ", d[0, 0, :8].tolist())
        print("   ref[0,0,:8]", ref_raw[0, 0, :8].tolist())
    if big:
        for prec, label in ((K.PREC_TF32, "tf32 tcgen05"), (K.PREC_FP32, "3xTF32 tcgen05"), (K.PREC_FP32_FFMA, "fp32 FFMA")):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for _ in range(2):
                K.conv_wgrad(dyc, xc, precision=prec, **kw)
            e0.record()
            for _ in range(5):
                K.conv_wgrad(dyc, xc, precision=prec, **kw)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            fl = 2.0 * nb * t_out * v * cin * cout * taps
            print(f"   {label}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s")


def run_case(name):
    if name in WG_CASES:
        return run_wg_case(name)
    import torch
    from fusion_gcn_b200 import ops as K
    from oracle import stages as S
    nb, t_in, v, cin, cout, taps, stride, transposed, use_bias, acc = CASES[name]
    pad = (taps - 1) // 2
    if transposed:
        t_out = t_in
        t_src = (t_in + 2 * pad - taps) // stride + 1
    else:
        t_out = (t_in + 2 * pad - taps) // stride + 1
        t_src = t_in
    g = torch.Generator().manual_seed(0)
    x = torch.randn(nb, t_src, v, cin, generator=g)
    w = torch.randn(cout, taps, cin, generator=g) * 0.1
    b = torch.randn(cout, generator=g) if use_bias else None
    base = torch.randn(nb, t_out, v, cout, generator=g) if acc else None
    kw = dict(t_out=t_out, stride=stride, pad=pad, transposed=bool(transposed))
    xc, wc = x.cuda(), w.cuda()
    out = base.cuda().clone() if acc else None
    y = K.conv_fwd(xc, wc, None if b is None else b.cuda(), out=out, accumulate=bool(acc), precision=K.PREC_TF32, **kw)
    torch.cuda.synchronize()
    out3 = base.cuda().clone() if acc else None
    y3 = K.conv_fwd(xc, wc, None if b is None else b.cuda(), out=out3, accumulate=bool(acc), precision=K.PREC_FP32, **kw)
    torch.cuda.synchronize()
    big = nb * t_out * v * cout > 5e7

    def ref(xx, ww):
        if big:   # fp32 FFMA kernel as the reference for the big shapes
            r = K.conv_fwd(xx.float().cuda(), ww.float().cuda(), None if b is None else b.cuda(), precision=K.PREC_FP32_FFMA, **kw).double().cpu()
        else:
            r = S.conv_fwd(xx.double(), ww.double(), None if b is None else b.double(), **kw)
        return r + base.double() if acc else r

    def trunc(t_):
        return (t_.view(torch.int32) & ~0x1FFF).view(torch.float32)

    def rnd(t_):
        return ((t_.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)

    yd = y.double().cpu()
    res = {}
    for tag, (xx, ww) in {"raw": (x, w), "trunc": (trunc(x), trunc(w)), "round": (rnd(x), rnd(w))}.items():
        r = ref(xx, ww)
        res[tag] = ((yd - r).abs().max() / r.abs().max()).item()
        if tag == "raw":
            err = (yd - r).abs()
            rr = r
    e3 = ((y3.double().cpu() - rr).abs().max() / rr.abs().max()).item()
    print(f"[{name}] tf32: vs raw {res['raw']:.3e} | vs truncated-TF32 {res['trunc']:.3e} || 3xTF32 (fp32 mode): vs raw {e3:.3e}")
    if res["raw"] > 5e-3:
        flat = err.reshape(-1, cout)
        bad_rows = (flat.max(dim=1).values > 1e-2 * rr.abs().max()).nonzero().flatten()
        bad_cols = (flat.max(dim=0).values > 1e-2 * rr.abs().max()).nonzero().flatten()
        print(f"   bad rows {bad_rows.numel()}/{flat.shape[0]}: {bad_rows[:24].tolist()} ...")
        print(f"   bad cols {bad_cols.numel()}/{cout}: {bad_cols[:24].tolist()} ...")
        print("   y[0,:8]  ", yd.reshape(-1, cout)[0, :8].tolist())
        print("   ref[0,:8]", rr.reshape(-1, cout)[0, :8].tolist())
    if big:
        for prec, label in ((K.PREC_TF32, "tf32 tcgen05"), (K.PREC_FP32, "3xTF32 tcgen05"), (K.PREC_FP32_FFMA, "fp32 FFMA")):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for _ in range(2):
                K.conv_fwd(xc, wc, None, precision=prec, **kw)
            e0.record()
            for _ in range(5):
                K.conv_fwd(xc, wc, None, precision=prec, **kw)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            fl = 2.0 * nb * t_out * v * cin * cout * taps
            print(f"   {label}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s  ({(x.numel() + nb * t_out * v * cout) * 4 / ms / 1e6:.0f} GB/s algorithmic)")


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        run_case(sys.argv[2])
        sys.exit(0)
    names = sys.argv[1:] or (list(CASES) + list(WG_CASES))
    for n in names:
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", n], timeout=90, capture_output=True, text=True)
            sys.stdout.write(p.stdout)
            if p.returncode != 0:
                print(f"[{n}] FAILED rc={p.returncode}: {p.stderr[-1500:]}")
        except subprocess.TimeoutExpired:
            print(f"[{n}] TIMEOUT (deadlock?)")
        sys.stdout.flush()
