"""-m gpu: the drop-in launcher end to end (SURVEY 4.5 training-curve smoke, VERDICT r1 missing item 11).

``python -m fusion_gcn_b200.dropin -- -f cfg.yaml`` runs the reference's UNMODIFIED torch_src/main.py (TrainingSession: YAML
config, MultiModalDataset over synthetic .npy files, optimizer / scheduler from the YAML, train + validation epochs) with the
B200 unit, the fused optimizers and the prefetching loader installed; ``--no-dropin --reference-fp32`` runs the same script on
the reference's own modules (cuDNN / cuBLAS, TF32 off) on the same GPU.  The per-step training losses of the two runs must
agree: tightly at the first steps, within a slowly growing bound afterwards (two fp32 implementations of SGD training drift
apart; a wrong kernel or optimizer shows up as an O(1) gap at step 1 or 2).  Needs the reference copy under baseline/_ref
(baseline/install_ref.py) or /root/reference."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import ref_loader

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write_dataset(root, modality, shape, n_train=54, n_val=27):
    rng = np.random.default_rng(0)
    os.makedirs(root, exist_ok=True)
    for split, n in (("train", n_train), ("val", n_val)):
        np.save(os.path.join(root, f"{modality}_{split}_features.npy"), rng.standard_normal((n,) + shape).astype(np.float32))
        np.save(os.path.join(root, f"{split}_labels.npy"), np.arange(n) % 27)          # all 27 UTD-MHAD classes present


def _run(tmp, tag, model, optimizer_block, extra):
    cfg = os.path.join(tmp, f"{tag}.yaml")
    trace = os.path.join(tmp, f"{tag}_loss.txt")
    with open(cfg, "w") as fh:
        fh.write(f"""
input_data:
  - path: {tmp}/data
    loader: NumpyDatasetLoader
out_path: {tmp}/out_{tag}
model: {model}
dataset: UTD-MHAD
session_type: training
fixed_seed: 1
batch_size: 9
epochs: 2
{optimizer_block}
""")
    cmd = [sys.executable, "-m", "fusion_gcn_b200.dropin", "--reference", ref_loader.REFERENCE_ROOT, "--trace-loss", trace] + extra + \
          ["--", "-f", cfg, "--disable_logging", "--disable_checkpointing"]
    res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stderr[-3000:]
    return [float(v) for v in open(trace).read().split()]


SGD = "base_lr: 0.002\noptimizer: SGD\noptimizer_args:\n  momentum: 0.9\n  nesterov: true\n  weight_decay: 0.0001\nlr_scheduler: multistep\nlr_scheduler_args:\n  milestones: [1]"
ADAM = "base_lr: 0.0005\noptimizer: ADAM\noptimizer_args:\n  weight_decay: 0.01"


@pytest.mark.parametrize("model,opt", [("agcn", SGD), ("mmargcn", ADAM)], ids=["agcn-sgd", "mmargcn-adam"])
def test_training_curve_matches_the_reference_on_the_same_gpu(tmp_path, model, opt):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if not ref_loader.available():
        pytest.skip("reference copy not present (python baseline/install_ref.py)")
    tmp = str(tmp_path)
    if model == "mmargcn":       # the multimodal dispatcher with the thin rgb_patch_features wrapper: per-joint embeddings as channels
        _write_dataset(os.path.join(tmp, "data"), "rgb", (1, 32, 20, 16))
        opt = opt + "\nmode: rgb_patch_features\nmodel_args:\n  num_layers: 4"
    else:                        # models/agcn/agcn.py, the original 2s-AGCN copy (PA, l1..l10)
        _write_dataset(os.path.join(tmp, "data"), "skeleton", (1, 32, 20, 3))
    ref = _run(tmp, "ref", model, opt, ["--no-dropin", "--reference-fp32"])
    ours = _run(tmp, "ours", model, opt, [])
    assert len(ref) == len(ours) == 12 and all(np.isfinite(ours))                       # 2 epochs x 6 batches of 9 (drop_last)
    print(f"{model}: reference {['%.4f' % v for v in ref]}\n{model}: ours      {['%.4f' % v for v in ours]}")
    # same init, same first batch: the first two losses agree to fp32 rounding; afterwards two fp32 implementations of the same
    # tiny-batch training drift (the reference on CPU against the torch stage backend drifts to 1e-2 by step 8 as well)
    for i, (a, b) in enumerate(zip(ours, ref)):
        assert abs(a - b) <= (1e-4 if i < 2 else 5e-2) * max(1.0, abs(b)), (i, a, b)
    # the same session with its training step replayed as one CUDA graph per batch (--graph-step): same kernels, same order, the
    # classifier + loss through the fused head where the model has one -- the curve of the eager drop-in to rounding, then the same drift
    graphed = _run(tmp, "ours_graph", model, opt, ["--graph-step"])
    print(f"{model}: graphed   {['%.4f' % v for v in graphed]}")
    assert len(graphed) == len(ours) and all(np.isfinite(graphed))
    for i, (a, b) in enumerate(zip(graphed, ours)):
        assert abs(a - b) <= (1e-4 if i < 2 else 5e-2) * max(1.0, abs(b)), (i, a, b)
