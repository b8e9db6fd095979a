"""dgl.function stand-ins (see package docstring). Test infrastructure only."""
import torch


class _BuiltinEdge:
    def __init__(self, f):
        self.f = f

    def __call__(self, g):
        return self.f(g)


def u_add_v(a, b, out):
    return _BuiltinEdge(lambda g: {out: g.ndata[a][g._src] + g.ndata[b][g._dst]})


def v_sub_u(a, b, out):
    return _BuiltinEdge(lambda g: {out: g.ndata[a][g._dst] - g.ndata[b][g._src]})


def u_mul_e(a, w, out):
    return lambda g: {out: g.ndata[a][g._src] * g.edata[w]}


def copy_e(w, out):
    return lambda g: {out: g.edata[w]}


def sum(msg, out):  # noqa: A001 - mirrors dgl.function.sum
    def red(g, m):
        v = m[msg]
        acc = torch.zeros((g.num_nodes(),) + tuple(v.shape[1:]), dtype=v.dtype)
        return {out: acc.index_add(0, g._dst, v)}
    return red
