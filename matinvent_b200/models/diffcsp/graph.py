"""Batch topology for the CSPNet kernels: crystals are disjoint graphs laid out back to back.

Replaces the per-forward `block_diag` + `nonzero` of CSPNet.gen_edges (models/diffcsp/cspnet.py:236-257):
the edge list depends only on `num_atoms`, so it is built once per batch by `mi_fc_edges` and reused by
every one of the 2 000 score-network evaluations of a sampling run.  All index arrays are int32.
"""
import torch

from ... import ops


class CrystalGraph:
    """Fully-connected ('fc') topology in CSR form over the source node.

    node_off [B+1], edge_off [B+1]: prefix sums of n_b and n_b^2
    edge_src / edge_dst / edge_graph [E]; seg_ptr [N+1] (edges of node i are contiguous)
    dst_ptr [N+1] + dst_perm [E]: the same edges grouped by destination (backward of the h_j gather)
    node_graph [N]
    """

    def __init__(self, num_atoms, device):
        na = torch.as_tensor(num_atoms).detach().to("cpu", torch.int64).reshape(-1)
        if na.numel() == 0:
            raise ValueError("empty batch")
        if int(na.min()) < 1:
            raise ValueError("every crystal needs at least one atom")
        self.device = torch.device(device)
        self.num_atoms_cpu = na
        self.B = int(na.numel())
        self.N = int(na.sum())
        self.E = int((na * na).sum())
        self.max_atoms = int(na.max())
        z = torch.zeros(1, dtype=torch.int64)
        i32 = dict(dtype=torch.int32, device=self.device)
        self.node_off = torch.cat([z, torch.cumsum(na, 0)]).to(**i32)
        self.edge_off = torch.cat([z, torch.cumsum(na * na, 0)]).to(**i32)
        self.num_atoms = na.to(self.device)                     # int64, reference-facing
        self.edge_src = torch.empty(self.E, **i32)
        self.edge_dst = torch.empty(self.E, **i32)
        self.edge_graph = torch.empty(self.E, **i32)
        self.seg_ptr = torch.empty(self.N + 1, **i32)
        self.dst_ptr = torch.empty(self.N + 1, **i32)
        self.dst_perm = torch.empty(self.E, **i32)
        self.node_graph = torch.empty(self.N, **i32)
        self.cell_off = None                                   # 'knn' only
        ops.fc_edges(self.node_off, self.edge_off, self.B, self.N, self.E, self.edge_src, self.edge_dst,
                     self.edge_graph, self.seg_ptr, self.dst_ptr, self.dst_perm, self.node_graph)
        self.edge_w = self.mean_weights(self.E)

    def mean_weights(self, E):
        """edge_w[e] = 1 / (edges of the source node of e): the weights of the scatter-mean fused into the second
        per-edge GEMM's epilogue (index table of the batch topology, built with it)"""
        deg = (self.seg_ptr[1:] - self.seg_ptr[:-1]).clamp(min=1).to(torch.float32)
        return (1.0 / deg)[self.edge_src[:E].long()].contiguous()

    @property
    def node2graph(self):
        """int64 PyG-style `batch` vector (reference-facing)."""
        return self.node_graph.to(torch.int64)

    def key(self):
        return tuple(self.num_atoms_cpu.tolist())
