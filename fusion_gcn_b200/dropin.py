"""Drop-in installer / launcher: runs the UNMODIFIED reference (main.py, YAML configs, fusion wrappers) on top of the
B200 unit.

The reference discovers its model with ``import_model(cfg.model)`` -> ``models.<name>.<name>.Model``
(torch_src/session/session.py:50, util/dynamic_import.py:29-38) and every multimodal wrapper builds the backbone as
``agcn.Model(...)`` through the module attribute (torch_src/models/mmargcn/early_fusion_models.py:18,36,71,108,144,187,
228,265; late_fusion_models.py:19,55; rgb_feature_models.py:22,43,89).  The reference tree is read-only, so the drop-in
is a rebinding of those module attributes:

    models.mmargcn.agcn.{TemporalConv, SpatialGraphConv, SpatialTemporalConv, Model}  -> fusion_gcn_b200.modules.*
    models.agcn.agcn.{unit_tcn, unit_gcn, TCN_GCN_unit, Model}                        -> fusion_gcn_b200.modules_original.*
    models.mmargcn.{graph_convolution, gcn}.{AGCNGraphConvolution, STGCNGraphConvolution} -> fusion_gcn_b200.graphconv.*

Usage (on a box that has both the reference checkout and a B200):

    python -m fusion_gcn_b200.dropin --reference /path/to/fusion-gcn [--precision fp32|tf32] -- -f config/<...>.yaml

which installs the import shims the reference needs on a current software stack, rebinds the
classes and then ``runpy``s ``torch_src/main.py`` with cwd = the reference root (its config code scans cwd for models and
datasets, torch_src/config.py:21-41).  YAML files, main.py and the session code stay byte-identical.
"""
import importlib
import os
import runpy
import sys
from unittest.mock import MagicMock

_ORIGINALS = {}          # (module name, attribute) -> reference object, for uninstall()

_MMARGCN_NAMES = ("TemporalConv", "SpatialGraphConv", "SpatialTemporalConv", "Model")
_ORIGINAL_NAMES = ("unit_tcn", "unit_gcn", "TCN_GCN_unit", "Model")


def install_shims(reference_root: str, host_stubs: bool = True) -> None:
    """Makes the reference importable on numpy >= 1.24 / a box without matplotlib, seaborn or ray."""
    import numpy as np
    for p in (os.path.join(reference_root, "torch_src"), reference_root):
        if p not in sys.path:
            sys.path.insert(0, p)
    names = ["matplotlib", "matplotlib.pyplot"]
    if host_stubs:
        names += ["seaborn", "ray", "ray.tune", "ray.tune.schedulers"]
    for name in names:
        try:
            importlib.import_module(name)
        except Exception:                                   # absent (or broken) optional dependency: stub it
            sys.modules.setdefault(name, MagicMock(name=name))
    if not hasattr(np, "int"):
        np.int = int                                        # util/graph.py:75,88
    if not hasattr(np, "float"):
        np.float = float                                    # util/graph.py:117,127
    # np.bool is deliberately left alone (numpy >= 2 defines it; overriding it breaks numpy.ma / scipy)


def install(reference_root: str, precision="fp32", host_stubs: bool = True, fused_optimizers: bool = False, prefetch: bool = False,
            recompute: bool = False) -> dict:
    """Rebinds the reference's unit / backbone classes to the B200 implementations.  Returns the patched modules.
    ``fused_optimizers``: the YAML optimizer names SGD / ADAM / ADAMW (torch_src/session_helper.py:48-53) build the multi-tensor
    classes of fusion_gcn_b200.optim.  ``prefetch``: the sessions' ``DataLoader`` (session/training.py:18-25, evaluation.py:20-24,
    debugging.py:15-21) becomes fusion_gcn_b200.pipeline.PrefetchLoader (same batches, pinned staging + asynchronous H2D)."""
    from . import modules, modules_original
    if not os.path.isfile(os.path.join(reference_root, "torch_src", "models", "mmargcn", "agcn.py")):
        raise FileNotFoundError(f"{reference_root} does not look like a fusion-gcn checkout (torch_src/models/mmargcn/agcn.py missing)")
    install_shims(reference_root, host_stubs)
    modules.set_default_precision(precision)
    modules.set_default_recompute(recompute)             # activation-recompute policy of the units the session builds
    ref_m = importlib.import_module("models.mmargcn.agcn")
    ref_o = importlib.import_module("models.agcn.agcn")
    for mod, names, impl in ((ref_m, _MMARGCN_NAMES, modules), (ref_o, _ORIGINAL_NAMES, modules_original)):
        for name in names:
            _ORIGINALS.setdefault((mod.__name__, name), getattr(mod, name))
            setattr(mod, name, getattr(impl, name))
    patched = {"models.mmargcn.agcn": ref_m, "models.agcn.agcn": ref_o}
    # 1-D graph convolutions of the IMU / late-fusion models (graph_convolution.py:12-113); gcn.py binds the names at import time
    from . import graphconv
    ref_g = importlib.import_module("models.mmargcn.graph_convolution")
    ref_gcn = importlib.import_module("models.mmargcn.gcn")
    for mod in (ref_g, ref_gcn):
        for name in ("AGCNGraphConvolution", "STGCNGraphConvolution"):
            _ORIGINALS.setdefault((mod.__name__, name), getattr(mod, name))
            setattr(mod, name, getattr(graphconv, name))
        patched[mod.__name__] = mod
    if fused_optimizers:
        from .optim import FUSED_OPTIMIZERS
        helper = importlib.import_module("session_helper")
        _ORIGINALS.setdefault((helper.__name__, "available_optimizers"), helper.available_optimizers)
        helper.available_optimizers = dict(helper.available_optimizers, **FUSED_OPTIMIZERS)
        patched["session_helper"] = helper
    if prefetch:
        import torch
        from .pipeline import PrefetchLoader
        device = "cuda" if torch.cuda.is_available() else "cpu"

        def loader(dataset, batch_size=1, shuffle=False, drop_last=False, **kw):
            return PrefetchLoader(dataset, batch_size, shuffle=shuffle, drop_last=drop_last, device=device)
        importlib.import_module("session_helper")      # first, as main.py does: session <-> session_helper import each other
        for name in ("session.training", "session.evaluation", "session.debugging"):
            mod = importlib.import_module(name)
            _ORIGINALS.setdefault((mod.__name__, "DataLoader"), mod.DataLoader)
            mod.DataLoader = loader
            patched[name] = mod
    return patched


def uninstall() -> None:
    """Restores the reference classes (used by the tests)."""
    for (mod_name, name), obj in _ORIGINALS.items():
        setattr(sys.modules[mod_name], name, obj)
    _ORIGINALS.clear()


def _graph_steps() -> None:
    """The sessions' training step (procedures/step.py:39-46: forward, loss, backward launched op by op from Python) replayed as ONE
    CUDA graph per step.  At the reference's own batch sizes (8 - 10) the eager step is launch bound: 20.8 ms against
    11.2 ms replayed at 8 sequences (profiles/r17_bench_b8.json).  ``DefaultStep.forward`` builds a ``graphed.GraphedStep`` lazily on
    the first training batch (BatchNorm buffers preserved across the warm-up), replays it for every batch of that shape and hands back
    (logits, loss); ``DefaultStep.backward`` then only re-attaches the graph's static gradients (the session's ``optimizer.zero_grad()``
    sets them to None).  Everything else -- evaluation, another batch shape (a shorter last batch), dict features of the fusion
    models, gradient accumulation, autocast -- takes the reference's own code path."""
    import torch
    from .graphed import GraphedStep
    importlib.import_module("session_helper")          # first, as main.py does: session <-> session_helper import each other
    step = importlib.import_module("session.procedures.step")
    inner_forward, inner_backward = step.DefaultStep.forward, step.DefaultStep.backward
    state = {"graph": None, "key": None, "pending": None, "disabled": False}

    def forward(self, model, loss_function, features, label, loss_quotient=1):
        state["pending"] = None
        usable = (not state["disabled"] and model.training and torch.is_grad_enabled() and isinstance(features, torch.Tensor)
                  and features.is_cuda and label.is_cuda and loss_quotient == 1 and not torch.is_autocast_enabled())
        if usable:
            key = (id(model), tuple(features.shape), tuple(label.shape), features.dtype, label.dtype)
            if state["graph"] is None:
                try:
                    state["graph"] = GraphedStep(model, loss_function, features, label, warmup=2, preserve_buffers=True)
                    state["key"] = key
                except Exception as exc:                 # noqa: BLE001 -- the eager step stands in for the rest of the session
                    state["disabled"] = True
                    model.zero_grad(set_to_none=True)
                    print(f"fusion_gcn_b200.dropin: CUDA-graph step disabled ({type(exc).__name__}: {exc})", file=sys.stderr)
            if state["graph"] is not None and key == state["key"]:
                gs = state["graph"]
                loss = gs(features, label)
                state["pending"] = gs
                return gs.logits, loss
        return inner_forward(self, model, loss_function, features, label, loss_quotient)

    def backward(self, loss):
        gs, state["pending"] = state["pending"], None
        if gs is not None and loss is gs.loss:
            gs.attach_grads()                            # the replay already ran the backward
        else:
            inner_backward(self, loss)
    step.DefaultStep.forward, step.DefaultStep.backward = forward, backward


def _trace_losses(path: str) -> None:
    """Appends every training-step loss (procedures/step.py:39-43) to ``path`` as text, one value per line."""
    importlib.import_module("session_helper")          # first, as main.py does: session <-> session_helper import each other
    step = importlib.import_module("session.procedures.step")
    inner = step.DefaultStep.forward

    def forward(self, model, loss_function, features, label, loss_quotient=1):
        y_pred, loss = inner(self, model, loss_function, features, label, loss_quotient)
        if model.training:
            with open(path, "a") as fh:
                fh.write(repr(float(loss.detach())) + "\n")
        return y_pred, loss
    step.DefaultStep.forward = forward


def main(argv=None) -> None:
    argv = list(sys.argv[1:] if argv is None else argv)
    rest = []
    if "--" in argv:
        i = argv.index("--")
        argv, rest = argv[:i], argv[i + 1:]
    import argparse
    ap = argparse.ArgumentParser(prog="python -m fusion_gcn_b200.dropin", description=__doc__.split("\n")[0])
    ap.add_argument("--reference", default=os.environ.get("FUSION_GCN_REFERENCE", "/root/reference"))
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16x3", "tf32", "fp32_ffma"])
    ap.add_argument("--recompute", action="store_true", help="units recompute theta / phi and the aggregated tensor in the backward (~40 %% less activation memory)")
    ap.add_argument("--graph-step", action="store_true", help="replay the training step (forward + loss + backward) as one CUDA graph per batch")
    ap.add_argument("--no-dropin", action="store_true", help="run the UNMODIFIED reference under the same import shims (A/B baseline)")
    ap.add_argument("--reference-fp32", action="store_true", help="with --no-dropin: disable cuDNN / cuBLAS TF32 so the reference is an fp32 oracle (SURVEY D9)")
    ap.add_argument("--no-fused-optimizers", action="store_true", help="keep torch.optim instead of fusion_gcn_b200.optim")
    ap.add_argument("--no-prefetch", action="store_true", help="keep torch's DataLoader instead of fusion_gcn_b200.pipeline.PrefetchLoader")
    ap.add_argument("--trace-loss", default=None, metavar="FILE", help="append every training-step loss to FILE")
    args = ap.parse_args(argv)
    root = os.path.abspath(args.reference)
    if args.no_dropin:
        if not os.path.isfile(os.path.join(root, "torch_src", "main.py")):
            raise FileNotFoundError(f"{root} does not look like a fusion-gcn checkout")
        install_shims(root)
        if args.reference_fp32:
            import torch
            torch.backends.cudnn.allow_tf32 = False
            torch.backends.cuda.matmul.allow_tf32 = False
    else:
        install(root, args.precision, fused_optimizers=not args.no_fused_optimizers, prefetch=not args.no_prefetch, recompute=args.recompute)
        from . import capi
        capi.lib()                                           # fail before training starts if the extension is not built
        if args.graph_step:
            _graph_steps()
    if args.trace_loss:
        _trace_losses(args.trace_loss)
    os.chdir(root)
    sys.argv = [os.path.join(root, "torch_src", "main.py")] + rest
    runpy.run_path(sys.argv[0], run_name="__main__")


if __name__ == "__main__":
    main()
