"""TEST INFRASTRUCTURE ONLY -- loader for the *real* reference modules.

Usable where /root/reference exists (the build container) or where ``baseline/_ref`` holds the byte copy made by
``baseline/install_ref.py`` (the GPU box).  It is used by
``oracle/make_golden.py`` to generate the committed fixtures under
``tests/golden/`` and by the ``not gpu`` tests that pin ``oracle/agcn_oracle.py``
against the reference itself.  Nothing on the GPU box may import the reference
(it does not exist there); callers must check :func:`available` first.

The shims follow SURVEY.md Appendix D: stub ``matplotlib`` (imported by
``util/graph.py:1``) and restore the ``np.int`` / ``np.float`` aliases that
``util/graph.py:75,88,117,127`` still use.  ``np.bool`` is never touched.
"""
import os
import sys
from unittest.mock import MagicMock

_TRAVELLING_COPY = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


def _pick_root():
    """/root/reference in the build container; on the GPU box the byte copy made by baseline/install_ref.py
    (git-ignored, shipped with the snapshot)."""
    env = os.environ.get("FUSION_GCN_REFERENCE")
    for cand in ([env] if env else []) + ["/root/reference", _TRAVELLING_COPY]:
        if os.path.isfile(os.path.join(cand, "torch_src", "models", "mmargcn", "agcn.py")):
            return cand
    return env or "/root/reference"


REFERENCE_ROOT = _pick_root()


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "torch_src", "models", "mmargcn", "agcn.py"))


_loaded = {}


def load():
    """Returns a dict with the reference's modules: ``agcn`` (mmargcn variant),
    ``Graph``, ``GraphPartitionStrategy``, ``ntu``, ``utd``, ``mmact`` constants
    and ``fusion`` (graph extension for IMU joints)."""
    if _loaded:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present at " + REFERENCE_ROOT)
    import numpy as np
    for p in (os.path.join(REFERENCE_ROOT, "torch_src"), REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    for name in ("matplotlib", "matplotlib.pyplot", "seaborn", "ray", "ray.tune", "ray.tune.schedulers"):
        sys.modules.setdefault(name, MagicMock(name=name))
    if not hasattr(np, "int"):
        np.int = int
    if not hasattr(np, "float"):
        np.float = float
    from util.graph import Graph
    from util.partition_strategy import GraphPartitionStrategy
    from models.mmargcn import agcn as ref_agcn
    from models.mmargcn import fusion as ref_fusion
    import datasets.ntu_rgb_d.constants as ntu
    import datasets.utd_mhad.constants as utd
    import datasets.mmact.constants as mmact
    _loaded.update(agcn=ref_agcn, Graph=Graph, GraphPartitionStrategy=GraphPartitionStrategy,
                   fusion=ref_fusion, ntu=ntu, utd=utd, mmact=mmact)
    return _loaded
