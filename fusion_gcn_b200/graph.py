"""Partition adjacency A_k for the 'spatial' strategy, and the skeleton edge lists of the datasets the
reference ships.

Restates util/partition_strategy.py:42-46 + util/graph.py:74-79,116-124 of the reference:
A[0] = I, A[1] = column-normalised reversed (centre -> limb) edges, A[2] = column-normalised
(limb -> centre) edges.  Columns with zero degree are written as explicit zeros (the reference leaves
them to uninitialised memory).  ``adjacency_from_graph`` accepts the reference's own
``util.graph.Graph`` objects (anything with ``.edges`` and ``.num_vertices``), which is what
``Session._build_model`` passes to ``Model(data_shape, num_classes, graph, ...)``.
"""
import numpy as np

# edge lists (i, j): i -> j points towards the centre joint
NTU_EDGES = [(0, 1), (1, 20), (2, 20), (3, 2), (4, 20), (5, 4), (6, 5), (7, 6), (8, 20), (9, 8), (10, 9), (11, 10),
             (12, 0), (13, 12), (14, 13), (15, 14), (16, 0), (17, 16), (18, 17), (19, 18), (21, 22), (22, 7), (23, 24),
             (24, 11)]                                                   # datasets/ntu_rgb_d/constants.py:108-134
NTU_CENTER = 20
UTD_EDGES = [(0, 1), (2, 1), (4, 1), (8, 1), (3, 2), (12, 3), (16, 3), (5, 4), (6, 5), (7, 6), (9, 8), (10, 9), (11, 10),
             (13, 12), (14, 13), (15, 14), (17, 16), (18, 17), (19, 18)]  # datasets/utd_mhad/constants.py:87-108
UTD_CENTER = 1
MMACT_EDGES = [(0, 1), (2, 1), (5, 1), (8, 1), (11, 1), (3, 2), (4, 3), (6, 5), (7, 6), (9, 8), (10, 9), (12, 11), (13, 12),
               (14, 0), (15, 0), (16, 14), (17, 15)]                      # datasets/mmact/constants.py:83-102 (COCO-18)
MMACT_CENTER = 1


class SkeletonGraph:
    """Minimal stand-in for the reference's util.graph.Graph (edges + num_vertices + center_joint)."""

    def __init__(self, edges, num_vertices=None, center_joint=0):
        self.edges = np.unique(np.asarray(edges, dtype=np.int64), axis=0)
        if self.edges.ndim != 2 or self.edges.shape[1] != 2 or (self.edges < 0).any():
            raise ValueError("edges must be an (E, 2) array of non-negative integers")
        nv = int(self.edges.max()) + 1
        if num_vertices is not None and num_vertices < nv:
            raise ValueError("num_vertices smaller than the largest vertex id")
        self.num_vertices = nv if num_vertices is None else int(num_vertices)
        self.center_joint = center_joint

    def with_new_edges(self, edges):
        return SkeletonGraph(np.vstack((self.edges, np.asarray(edges, dtype=np.int64))), center_joint=self.center_joint)


def partition_adjacency(edges, num_vertices=None, strategy="spatial"):
    if strategy == "distance":
        raise NotImplementedError("Distance strategy not implemented (as in the reference, util/partition_strategy.py:28)")
    if strategy not in ("spatial", "uniform"):
        raise ValueError("Unsupported partition strategy: " + str(strategy))
    e = np.unique(np.asarray(edges, dtype=np.int64), axis=0)
    v = int(e.max()) + 1 if num_vertices is None else int(num_vertices)
    inward = np.zeros((v, v), dtype=np.float64)
    inward[e[:, 0], e[:, 1]] = 1.0

    def col_norm(a):
        deg = a.sum(axis=0)
        scale = np.divide(1.0, deg, out=np.zeros_like(deg), where=deg > 0)
        return a * scale[None, :]

    if strategy == "uniform":
        return col_norm(np.maximum(inward, inward.T))[None]
    return np.stack([np.eye(v), col_norm(inward.T), col_norm(inward)])


def adjacency_from_graph(graph, strategy="spatial"):
    return partition_adjacency(graph.edges, graph.num_vertices, strategy)


def imu_fusion_graph(graph, num_imu_joints, mode="append_center", interconnect=False, **kwargs):
    """torch_src/models/mmargcn/fusion.py:65-89."""
    nv = graph.num_vertices
    new = []
    if mode == "append_center":
        centre = kwargs.get("center_joint", graph.center_joint)
        new += [(nv + i, centre) for i in range(num_imu_joints)]
    elif mode == "append_right":
        for i in range(num_imu_joints):
            new += [(nv + i, kwargs["right_wrist_joint"]), (nv + i, kwargs["right_hip_joint"])]
    else:
        raise ValueError("Unsupported imu_enhanced_mode: " + str(mode))
    if interconnect or kwargs.get("interconnect_imu_joints", False):       # the reference's keyword (fusion.py:84)
        new += [(nv + i, nv + j) for i in range(num_imu_joints) for j in range(i + 1, num_imu_joints)]
    return SkeletonGraph(np.vstack((np.asarray(graph.edges), np.asarray(new, dtype=np.int64))), center_joint=graph.center_joint)
