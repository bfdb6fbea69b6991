"""Periodic k-nearest-neighbour topology (edge_style='knn') rebuilt on the device every forward.

Replaces radius_graph_pbc + get_max_neighbors_mask (models/diffcsp/utils.py:335-601) and
reorder_symmetric_edges (models/diffcsp/cspnet.py:159-234): one CTA per crystal finds, per centre atom,
the candidates within the adaptive radius over the 27 neighbouring images, applies the reference's
"(K+1)-th nearest + 0.01" cap, keeps one direction of every pair and emits both directions grouped by
source node (the network is invariant to edge order; only the CSR grouping differs from the reference).
The edge count is data dependent, so this path reads it back once per forward and is not graph-captured.
"""
import torch

from ... import ops
from .graph import CrystalGraph


class KnnGraph(CrystalGraph):
    def __init__(self, num_atoms, device, max_neighbors):
        super().__init__(num_atoms, device)          # node arrays (+ unused fc edges)
        self.max_neighbors = int(max_neighbors)
        self.cap = 3 * (self.max_neighbors + 2)
        N = self.N
        i32 = dict(dtype=torch.int32, device=self.device)
        self.E_cap = N * self.cap
        self._dst_pad = torch.empty(N * self.cap, **i32)
        self._cell_pad = torch.empty(N * self.cap, 3, device=self.device)
        self._deg = torch.empty(N, **i32)
        self._overflow = torch.zeros(1, **i32)
        self._work = torch.empty(N + 1, **i32)
        self.edge_src = torch.zeros(self.E_cap, **i32)
        self.edge_dst = torch.zeros(self.E_cap, **i32)
        self.edge_graph = torch.zeros(self.E_cap, **i32)
        self.dst_perm = torch.zeros(self.E_cap, **i32)
        self.cell_off = torch.zeros(self.E_cap, 3, device=self.device)
        self.E = self.E_cap

    def rebuild(self, x, l, need_dst=True):
        ops.radius_graph_pbc(x, l, self.node_off, self.B, self.N, self.max_atoms, self.max_neighbors, self.cap,
                             self._dst_pad, self._cell_pad, self._deg, self._overflow)
        ops.compact_edges(self._deg, self.N, self.cap, self._dst_pad, self._cell_pad, self.node_graph, self.seg_ptr,
                          self.edge_src, self.edge_dst, self.edge_graph, self.cell_off, self.E_cap)
        tail = torch.stack([self.seg_ptr[self.N], self._overflow[0]]).tolist()      # one D2H sync
        if tail[1]:
            raise RuntimeError("knn neighbour list overflow: a node has more than %d symmetric edges" % self.cap)
        self.E = int(tail[0])
        self.edge_w = self.mean_weights(self.E)
        if need_dst:
            ops.build_dst_csr(self.seg_ptr, self.edge_dst, self.N, self.E_cap, self.dst_ptr, self.dst_perm, self._work)
        return self
