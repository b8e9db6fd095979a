"""Minimal pure-torch stand-in for the DGL API subset the reference model touches.

TEST INFRASTRUCTURE ONLY (lives under oracle/): it exists so that the UNMODIFIED
reference modules under /root/reference (alignn/models/alignn.py,
alignn/models/alignn_atomwise.py) can be imported in the authoring container --
where the real `dgl` wheel is absent and cannot be installed -- to generate the
golden vectors in tests/golden/ (see oracle/make_golden.py).

DGL itself is an un-vendored third-party dependency of the reference
(pins: dgl<=1.1.1 setup.py:24, dgl==2.1.0 environment.yml:237).  The semantics
restated here are DGL's published ones:

  fn.u_add_v(a,b,o)      : edata[o][e] = ndata[a][src[e]] + ndata[b][dst[e]]
  fn.v_sub_u(a,b,o)      : edata[o][e] = ndata[a][dst[e]] - ndata[b][src[e]]
  fn.u_mul_e(a,w,o)      : message m[e] = ndata[a][src[e]] * edata[w][e]
  fn.copy_e(w,o)         : message m[e] = edata[w][e]
  fn.sum(m,o)            : ndata[o][v] = sum_{e: dst[e]=v} m[e]   (0 if no in-edge)
  g.line_graph(shared)   : node i == edge i of g; edge (i->j) iff dst(i)==src(j),
                           i != j, backtracking pairs kept; emitted in (i, j) order
  dgl.batch              : concatenation with id offsets, order preserved
  dgl.reverse            : swap src/dst, keep edge ids (and edata if asked)
  AvgPooling/SumPooling  : per-graph mean/sum over batch_num_nodes segments

Nothing in the product package (alignn_b200/) imports this.
"""
import torch
from . import function  # noqa: F401
from . import nn  # noqa: F401
from . import data  # noqa: F401


class _Frame(dict):
    pass


class _EdgeBatch:
    def __init__(self, g):
        s, d = g._src, g._dst
        self.src = {k: v[s] for k, v in g.ndata.items()}
        self.dst = {k: v[d] for k, v in g.ndata.items()}
        self.data = dict(g.edata)


class DGLGraph:
    def __init__(self, src, dst, num_nodes, bnn=None, bne=None):
        self._src = torch.as_tensor(src, dtype=torch.int64)
        self._dst = torch.as_tensor(dst, dtype=torch.int64)
        self._n = int(num_nodes)
        self.ndata = _Frame()
        self.edata = _Frame()
        self._bnn = torch.tensor([self._n]) if bnn is None else bnn
        self._bne = torch.tensor([len(self._src)]) if bne is None else bne

    # -- structure ------------------------------------------------------
    def edges(self):
        return self._src, self._dst

    def num_nodes(self):
        return self._n

    def num_edges(self):
        return int(self._src.numel())

    number_of_nodes = num_nodes
    number_of_edges = num_edges

    @property
    def batch_size(self):
        return int(self._bnn.numel())

    def batch_num_nodes(self):
        return self._bnn

    def batch_num_edges(self):
        return self._bne

    @property
    def device(self):
        return self._src.device

    def to(self, device):
        return self

    def local_var(self):
        g = DGLGraph(self._src, self._dst, self._n, self._bnn, self._bne)
        g.ndata.update(self.ndata)
        g.edata.update(self.edata)
        return g

    def __len__(self):  # ALIGNNAtomWise calls len(g) on the input tuple only
        raise TypeError("DGLGraph has no len()")

    # -- message passing ------------------------------------------------
    def apply_edges(self, func):
        out = func(self) if isinstance(func, function._BuiltinEdge) else func(_EdgeBatch(self))
        self.edata.update(out)

    def update_all(self, msg, red):
        m = msg(self)
        self.ndata.update(red(self, m))

    def line_graph(self, backtracking=True, shared=False):
        assert backtracking
        E = self.num_edges()
        src, dst = self._src, self._dst
        # out-edge lists of g keyed by source node, edge ids ascending
        order = torch.argsort(src, stable=True)
        counts = torch.bincount(src, minlength=self._n)
        ptr = torch.zeros(self._n + 1, dtype=torch.int64)
        ptr[1:] = torch.cumsum(counts, 0)
        deg = counts[dst]                       # number of successors j of edge i
        li = torch.repeat_interleave(torch.arange(E), deg)
        start = torch.repeat_interleave(ptr[dst], deg)
        off = torch.arange(li.numel()) - torch.repeat_interleave(
            torch.cumsum(deg, 0) - deg, deg)
        lj = order[start + off]
        keep = li != lj
        li, lj = li[keep], lj[keep]
        # batch bookkeeping
        bne = self._bne
        eoff = torch.cumsum(bne, 0)
        gid = torch.bucketize(li, eoff, right=True)
        lbne = torch.bincount(gid, minlength=bne.numel())
        lg = DGLGraph(li, lj, E, bne.clone(), lbne)
        if shared:
            lg.ndata.update(self.edata)
        return lg


def graph(data, num_nodes=None):
    src, dst = data
    src = torch.as_tensor(src, dtype=torch.int64)
    dst = torch.as_tensor(dst, dtype=torch.int64)
    if num_nodes is None:
        num_nodes = int(max(src.max(), dst.max())) + 1 if src.numel() else 0
    return DGLGraph(src, dst, num_nodes)


def batch(graphs):
    noff, srcs, dsts = 0, [], []
    for g in graphs:
        srcs.append(g._src + noff)
        dsts.append(g._dst + noff)
        noff += g._n
    bg = DGLGraph(torch.cat(srcs), torch.cat(dsts), noff,
                  torch.cat([g._bnn for g in graphs]),
                  torch.cat([g._bne for g in graphs]))
    for k in graphs[0].ndata:
        bg.ndata[k] = torch.cat([g.ndata[k] for g in graphs], 0)
    for k in graphs[0].edata:
        bg.edata[k] = torch.cat([g.edata[k] for g in graphs], 0)
    return bg


def unbatch(g):
    out, no, eo = [], 0, 0
    for n, e in zip(g._bnn.tolist(), g._bne.tolist()):
        h = DGLGraph(g._src[eo:eo + e] - no, g._dst[eo:eo + e] - no, n)
        for k, v in g.ndata.items():
            h.ndata[k] = v[no:no + n]
        for k, v in g.edata.items():
            h.edata[k] = v[eo:eo + e]
        out.append(h)
        no += n
        eo += e
    return out


def reverse(g, copy_ndata=True, copy_edata=False):
    r = DGLGraph(g._dst, g._src, g._n, g._bnn, g._bne)
    if copy_ndata:
        r.ndata.update(g.ndata)
    if copy_edata:
        r.edata.update(g.edata)
    return r


def radius_graph(x, r):
    """Non-periodic radius graph: edge u->v for every ordered pair with |x_u-x_v| < r, u != v."""
    with torch.no_grad():
        d = torch.cdist(x, x)
        mask = (d < r) & ~torch.eye(x.shape[0], dtype=torch.bool)
        dst, src = torch.nonzero(mask, as_tuple=True)   # grouped by destination
    return DGLGraph(src, dst, x.shape[0])
