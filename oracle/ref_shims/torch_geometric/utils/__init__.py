import torch


def to_dense_batch(x, batch=None, fill_value=0.0, max_num_nodes=None, batch_size=None):
    """Dense (B, Nmax, *) tensor + validity mask from a PyG-style (x, batch) pair."""
    if batch is None:
        batch = x.new_zeros(x.size(0), dtype=torch.long)
    if batch_size is None:
        batch_size = int(batch.max()) + 1 if batch.numel() else 0
    num_nodes = torch.zeros(batch_size, dtype=torch.long, device=batch.device).scatter_add_(
        0, batch, torch.ones_like(batch))
    cum = torch.cat([num_nodes.new_zeros(1), num_nodes.cumsum(0)])
    if max_num_nodes is None:
        max_num_nodes = int(num_nodes.max()) if batch_size else 0
    idx = torch.arange(batch.size(0), device=batch.device) - cum[batch] + batch * max_num_nodes
    size = [batch_size * max_num_nodes] + list(x.size())[1:]
    out = x.new_full(size, fill_value)
    out[idx] = x
    out = out.view([batch_size, max_num_nodes] + list(x.size())[1:])
    mask = torch.zeros(batch_size * max_num_nodes, dtype=torch.bool, device=batch.device)
    mask[idx] = True
    return out, mask.view(batch_size, max_num_nodes)


def to_dense_adj(edge_index, batch=None, edge_attr=None, max_num_nodes=None, batch_size=None):
    """Dense (B, Nmax, Nmax, *) adjacency: out[b, i, j] += edge_attr[e] for edge e = (i -> j) of graph b (node indices made
    local to their graph).  Used by the relation costs (models/clg/const.py), which only sum it over (i, j)."""
    if batch is None:
        n = int(edge_index.max()) + 1 if edge_index.numel() else 0
        batch = edge_index.new_zeros(n)
    if batch_size is None:
        batch_size = int(batch.max()) + 1 if batch.numel() else 1
    num_nodes = torch.zeros(batch_size, dtype=torch.long, device=batch.device).scatter_add_(0, batch, torch.ones_like(batch))
    cum = torch.cat([num_nodes.new_zeros(1), num_nodes.cumsum(0)])
    if max_num_nodes is None:
        max_num_nodes = int(num_nodes.max()) if batch_size else 0
    if edge_attr is None:
        edge_attr = torch.ones(edge_index.size(1), dtype=torch.float, device=batch.device)
    b = batch[edge_index[0]]
    i = edge_index[0] - cum[b]
    j = edge_index[1] - cum[b]
    size = [batch_size * max_num_nodes * max_num_nodes] + list(edge_attr.size())[1:]
    out = edge_attr.new_zeros(size)
    out = out.index_add(0, b * max_num_nodes * max_num_nodes + i * max_num_nodes + j, edge_attr)
    return out.view([batch_size, max_num_nodes, max_num_nodes] + list(edge_attr.size())[1:])
