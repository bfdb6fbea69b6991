"""Shim: PyG dense_to_sparse = row-major nonzero of a 2-D adjacency."""


def dense_to_sparse(adj):
    idx = adj.nonzero().t().contiguous()
    return idx, adj[idx[0], idx[1]]
