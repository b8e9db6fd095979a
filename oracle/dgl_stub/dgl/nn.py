"""dgl.nn stand-ins: per-graph segment pooling. Test infrastructure only."""
import torch


def _seg(g, x, mean):
    bnn = g.batch_num_nodes()
    gid = torch.repeat_interleave(torch.arange(bnn.numel()), bnn)
    out = torch.zeros((bnn.numel(),) + tuple(x.shape[1:]), dtype=x.dtype).index_add(0, gid, x)
    if mean:
        out = out / bnn.to(x.dtype).clamp(min=1).reshape(-1, *([1] * (x.dim() - 1)))
    return out


class AvgPooling(torch.nn.Module):
    def forward(self, g, x):
        return _seg(g, x, True)


class SumPooling(torch.nn.Module):
    def forward(self, g, x):
        return _seg(g, x, False)
