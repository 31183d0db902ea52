"""Drop-in modules for the original 2s-AGCN copy in the reference (torch_src/models/agcn/agcn.py).

Same maths as ``modules.py``; the differences mirrored here are the reference's own (agcn.py:55-63,97,136-162):
class names ``unit_tcn`` / ``unit_gcn`` / ``TCN_GCN_unit``, the learned adjacency is called ``PA``, the fixed
adjacency ``A`` is a plain tensor attribute (not in the state dict) that follows the input's device, layers are
``l1`` .. ``l10``, ``Model`` takes ``data_shape["skeleton"]`` and an optional ``adjacency_matrix=`` kwarg.
"""
import numpy as np
import torch
import torch.nn as nn

from . import modules as M
from .graph import adjacency_from_graph


class unit_tcn(M.TemporalConv):
    def __init__(self, in_channels, out_channels, kernel_size=9, stride=1):
        super().__init__(in_channels, out_channels, kernel_size=kernel_size, stride=stride)
        self.relu = nn.ReLU()         # present (unused) in the reference, agcn.py:46


class unit_gcn(M.SpatialGraphConv):
    def __init__(self, in_channels, out_channels, A, coff_embedding=4, num_subset=3):
        super().__init__(in_channels, out_channels, A, coff_embedding=coff_embedding, num_subsets=num_subset)
        fixed = self.adj_a.detach().clone()
        learned = self.adj_b.detach().clone()
        del self.adj_b
        del self._buffers["adj_a"]
        self.PA = nn.Parameter(learned)
        self.A = fixed                 # plain attribute, like the reference's Variable (agcn.py:62)
        self.inter_c = self.inter_channels
        self.num_subset = num_subset

    def _adj_fixed(self, x):
        if self.A.device != x.device:  # the reference copies A to the device every forward (agcn.py:97); cache it instead
            self.A = self.A.to(x.device)
        return self.A

    def _adj_learned(self):
        return self.PA


class TCN_GCN_unit(M.SpatialTemporalConv):
    _gcn_cls = unit_gcn
    _tcn_cls = unit_tcn

    def __init__(self, in_channels, out_channels, A, stride=1, residual=True):
        super().__init__(in_channels, out_channels, A, stride=stride, residual=residual)


class Model(M.Model):
    _unit_cls = TCN_GCN_unit

    def __init__(self, data_shape, num_classes, graph, **kwargs):
        shape = data_shape["skeleton"] if isinstance(data_shape, dict) else data_shape
        adj = kwargs.get("adjacency_matrix", None)
        if adj is None:
            adj = adjacency_from_graph(graph)
        # the reference reads nothing else from **kwargs (agcn.py:137-146): unknown model_args from a YAML (``mode`` ...) are
        # ignored there, so they are ignored here too; the size options of the mmargcn variant are accepted as an extension
        known = {k: v for k, v in kwargs.items() if k in ("num_layers", "start_feature_size", "without_fc", "dropout")}
        super().__init__(shape, num_classes, graph, adjacency_matrix=np.asarray(adj), **known)

    def _register_layers(self):
        for layer_idx, layer in enumerate(self.layers):
            setattr(self, f"l{layer_idx + 1}", layer)
