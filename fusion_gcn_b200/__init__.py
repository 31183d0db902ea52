"""fusion_gcn_b200 -- B200-native (sm_100a) implementation of fusion-gcn's AGCN / MMARGCN hot path.

Public surface mirrors the reference's torch_src/models/mmargcn/agcn.py and torch_src/models/agcn/agcn.py:
  modules.{TemporalConv, SpatialGraphConv, SpatialTemporalConv, Model}
  modules_original.{unit_tcn, unit_gcn, TCN_GCN_unit, Model}
The arithmetic lives in libagcn_b200.so (C ABI: include/agcn_b200.h); there is no CPU fallback.
"""
from . import capi, graph  # noqa: F401
from .modules import Model, SpatialGraphConv, SpatialTemporalConv, TemporalConv, set_precision, set_recompute, set_sync_batchnorm  # noqa: F401

__all__ = ["Model", "SpatialGraphConv", "SpatialTemporalConv", "TemporalConv", "set_precision", "set_recompute", "set_sync_batchnorm", "capi", "graph"]
