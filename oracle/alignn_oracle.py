"""CPU oracle: DGL-free restatement of the reference's edge-gated conv hot path.

THIS IS TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it; the product package
(alignn_b200/) never does.  It is the checker, never the thing shipped.

Every function cites the reference file:line (relative to /root/reference) it
follows.  The arithmetic that the reference delegates to DGL (an un-vendored
dependency, pins dgl<=1.1.1 setup.py:24 / dgl==2.1.0 environment.yml:237) is
restated from DGL's published semantics:
    u_add_v        out[e] = a[src[e]] + b[dst[e]]
    u_mul_e -> sum out[v] = sum_{e: dst[e]=v} a[src[e]] * w[e]   (0 if isolated)
    copy_e  -> sum out[v] = sum_{e: dst[e]=v} w[e]
and executed the way DGL's CPU backend does it (gather -> multiply -> index_add).

Parity pinning: the reference holds no golden vectors for this path
(SURVEY.md section 8c).  This oracle is pinned against outputs of the UNMODIFIED
reference modules (alignn/models/alignn.py, alignn/models/alignn_atomwise.py)
executed in the authoring container on top of oracle/dgl_stub (a stand-in for the
absent DGL wheel); the vectors are committed under tests/golden/ with the script
that made them (oracle/make_golden.py).  What stays unpinned is DGL's own kernels
(not installable offline) -- their semantics are the published ones above.

Works in fp32 and fp64; pure torch on CPU.
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn
from torch.nn import functional as F


# ----------------------------------------------------------------------------
# graph container (what dgl.DGLGraph provides to the path)
# ----------------------------------------------------------------------------
class OGraph:
    """src/dst int64 arrays + per-graph node/edge counts + feature dicts."""

    def __init__(self, src, dst, num_nodes, bnn=None, bne=None):
        self.src = torch.as_tensor(src, dtype=torch.int64)
        self.dst = torch.as_tensor(dst, dtype=torch.int64)
        self.n = int(num_nodes)
        self.bnn = torch.tensor([self.n]) if bnn is None else torch.as_tensor(bnn, dtype=torch.int64)
        self.bne = torch.tensor([self.src.numel()]) if bne is None else torch.as_tensor(bne, dtype=torch.int64)
        self.ndata, self.edata = {}, {}

    def num_nodes(self):
        return self.n

    def num_edges(self):
        return int(self.src.numel())


def line_graph(g: OGraph) -> OGraph:
    """g.line_graph(shared=True) as used at alignn/graphs.py:588.

    Node i of L(g) == edge i of g; edge (i -> j) iff dst(i) == src(j) and i != j
    (backtracking pairs kept).  Emitted sorted by (i, j).  L(g)'s edge order is
    not observable in any model output (z is consumed, never returned).
    """
    src, dst = g.src.numpy(), g.dst.numpy()
    E = src.shape[0]
    order = np.argsort(src, kind="stable")
    counts = np.bincount(src, minlength=g.n)
    ptr = np.zeros(g.n + 1, dtype=np.int64)
    ptr[1:] = np.cumsum(counts)
    deg = counts[dst]
    li = np.repeat(np.arange(E), deg)
    start = np.repeat(ptr[dst], deg)
    off = np.arange(li.shape[0]) - np.repeat(np.cumsum(deg) - deg, deg)
    lj = order[start + off]
    keep = li != lj
    li, lj = li[keep], lj[keep]
    eoff = np.cumsum(g.bne.numpy())
    gid = np.searchsorted(eoff, li, side="right")
    lbne = np.bincount(gid, minlength=len(eoff))
    lg = OGraph(li, lj, E, g.bne.clone(), lbne)
    lg.ndata = dict(g.edata)
    return lg


def batch(graphs) -> OGraph:
    """dgl.batch as used at alignn/lmdb_dataset.py:93-94: concat with id offsets."""
    noff, s, d = 0, [], []
    for g in graphs:
        s.append(g.src + noff)
        d.append(g.dst + noff)
        noff += g.n
    bg = OGraph(torch.cat(s), torch.cat(d), noff,
                torch.cat([g.bnn for g in graphs]), torch.cat([g.bne for g in graphs]))
    for k in graphs[0].ndata:
        bg.ndata[k] = torch.cat([g.ndata[k] for g in graphs], 0)
    for k in graphs[0].edata:
        bg.edata[k] = torch.cat([g.edata[k] for g in graphs], 0)
    return bg


def csr_by_key(key: np.ndarray, n: int):
    """Stable counting sort of edge ids by `key` (dst -> in-CSR, src -> out-CSR).

    Returns (ptr[n+1] int32, eid[E] int32).  Integer-exact reference for the
    product's index builder (bit-exact comparison in tests).
    """
    key = np.asarray(key, dtype=np.int64)
    eid = np.argsort(key, kind="stable").astype(np.int32)
    ptr = np.zeros(n + 1, dtype=np.int64)
    ptr[1:] = np.cumsum(np.bincount(key, minlength=n))
    return ptr.astype(np.int32), eid


def bond_cosines(r: torch.Tensor, lsrc: torch.Tensor, ldst: torch.Tensor) -> torch.Tensor:
    """compute_bond_cosines, alignn/graphs.py:847-864."""
    r1 = -r[lsrc]
    r2 = r[ldst]
    c = torch.sum(r1 * r2, dim=1) / (torch.norm(r1, dim=1) * torch.norm(r2, dim=1))
    return torch.clamp(c, -1, 1)


# ----------------------------------------------------------------------------
# model pieces
# ----------------------------------------------------------------------------
class RBFExpansion(nn.Module):
    """alignn/models/utils.py:11-44 (gamma = 1/lengthscale in the default branch, :30-34)."""

    def __init__(self, vmin=0.0, vmax=8.0, bins=40, lengthscale=None):
        super().__init__()
        self.vmin, self.vmax, self.bins = vmin, vmax, bins
        self.register_buffer("centers", torch.linspace(vmin, vmax, bins))
        if lengthscale is None:
            self.lengthscale = np.diff(self.centers).mean()
            self.gamma = 1 / self.lengthscale
        else:
            self.lengthscale = lengthscale
            self.gamma = 1 / (lengthscale ** 2)

    def forward(self, distance):
        return torch.exp(-self.gamma * (distance.unsqueeze(1) - self.centers) ** 2)


def _norm(kind, d):
    # BatchNorm1d: alignn/models/alignn.py:72,76,178 ; LayerNorm: alignn_atomwise.py:151,155, utils.py:285
    return nn.BatchNorm1d(d) if kind == "batchnorm" else nn.LayerNorm(d)


class MLPLayer(nn.Module):
    """Linear -> norm -> SiLU. alignn/models/alignn.py:170-184 (BN), models/utils.py:277-292 (LN)."""

    def __init__(self, fin, fout, norm="batchnorm"):
        super().__init__()
        self.layer = nn.Sequential(nn.Linear(fin, fout), _norm(norm, fout), nn.SiLU())

    def forward(self, x):
        return self.layer(x)


class EdgeGatedGraphConv(nn.Module):
    """alignn/models/alignn.py:48-129 (BatchNorm) / alignn_atomwise.py:127-208 (LayerNorm)."""

    def __init__(self, input_features, output_features, residual=True, norm="batchnorm"):
        super().__init__()
        self.residual = residual
        self.src_gate = nn.Linear(input_features, output_features)      # :68
        self.dst_gate = nn.Linear(input_features, output_features)      # :69
        self.edge_gate = nn.Linear(input_features, output_features)     # :70
        self.bn_edges = _norm(norm, output_features)                    # :71
        self.src_update = nn.Linear(input_features, output_features)    # :73
        self.dst_update = nn.Linear(input_features, output_features)    # :74
        self.bn_nodes = _norm(norm, output_features)                    # :75

    def forward(self, g: OGraph, node_feats, edge_feats):
        src, dst = g.src, g.dst
        e_src = self.src_gate(node_feats)                               # :98
        e_dst = self.dst_gate(node_feats)                               # :99
        m = e_src[src] + e_dst[dst] + self.edge_gate(edge_feats)       # :100-101  u_add_v
        sigma = torch.sigmoid(m)                                        # :103
        Bh = self.dst_update(node_feats)                                # :104
        zeros = torch.zeros_like(Bh)
        sum_sigma_h = zeros.index_add(0, dst, Bh[src] * sigma)          # :105-107  u_mul_e -> sum
        sum_sigma = zeros.index_add(0, dst, sigma)                      # :108      copy_e -> sum
        h = sum_sigma_h / (sum_sigma + 1e-6)                            # :109
        x = self.src_update(node_feats) + h                             # :110
        x = F.silu(self.bn_nodes(x))                                    # :122
        y = F.silu(self.bn_edges(m))                                    # :123
        if self.residual:                                               # :125-127
            x = node_feats + x
            y = edge_feats + y
        return x, y


class ALIGNNConv(nn.Module):
    """alignn/models/alignn.py:132-167."""

    def __init__(self, fin, fout, norm="batchnorm"):
        super().__init__()
        self.node_update = EdgeGatedGraphConv(fin, fout, norm=norm)
        self.edge_update = EdgeGatedGraphConv(fout, fout, norm=norm)

    def forward(self, g, lg, x, y, z):
        x, m = self.node_update(g, x, y)         # :162
        y, z = self.edge_update(lg, m, z)        # :165  (L(g) node id == g edge id)
        return x, y, z


def avg_pool(g: OGraph, x):
    """dgl.nn.AvgPooling (alignn.py:242,325): per-graph mean over batch_num_nodes segments."""
    gid = torch.repeat_interleave(torch.arange(g.bnn.numel()), g.bnn)
    s = torch.zeros(g.bnn.numel(), x.shape[1], dtype=x.dtype).index_add(0, gid, x)
    return s / g.bnn.to(x.dtype).unsqueeze(1)


class ALIGNN(nn.Module):
    """alignn/models/alignn.py:187-349 (default branch: no extra_features).

    `norm="layernorm"` gives the conv/MLP stack of ALIGNNAtomWise
    (alignn_atomwise.py:272-333, energy-only path :364-470) with the same names.
    """

    def __init__(self, alignn_layers=4, gcn_layers=4, atom_input_features=92,
                 edge_input_features=80, triplet_input_features=40,
                 embedding_features=64, hidden_features=256, output_features=1,
                 link="identity", classification=False, num_classes=2, norm="batchnorm"):
        super().__init__()
        self.classification = classification
        self.atom_embedding = MLPLayer(atom_input_features, hidden_features, norm)       # :201
        self.edge_embedding = nn.Sequential(                                             # :205
            RBFExpansion(0, 8.0, edge_input_features),
            MLPLayer(edge_input_features, embedding_features, norm),
            MLPLayer(embedding_features, hidden_features, norm))
        self.angle_embedding = nn.Sequential(                                            # :214
            RBFExpansion(-1, 1.0, triplet_input_features),
            MLPLayer(triplet_input_features, embedding_features, norm),
            MLPLayer(embedding_features, hidden_features, norm))
        self.alignn_layers = nn.ModuleList(
            [ALIGNNConv(hidden_features, hidden_features, norm) for _ in range(alignn_layers)])
        self.gcn_layers = nn.ModuleList(
            [EdgeGatedGraphConv(hidden_features, hidden_features, norm=norm) for _ in range(gcn_layers)])
        if classification:
            self.fc = nn.Linear(hidden_features, num_classes)
            self.softmax = nn.LogSoftmax(dim=1)
        else:
            self.fc = nn.Linear(hidden_features, output_features)
        self.link_name = link
        if link == "log":                                                                # :273-278
            self.fc.bias.data = torch.tensor(np.log(0.7), dtype=torch.float)

    def conv_stack(self, g, lg, x, y, z):
        for layer in self.alignn_layers:          # :317-318
            x, y, z = layer(g, lg, x, y, z)
        for layer in self.gcn_layers:             # :321-322
            x, y = layer(g, x, y)
        return x, y

    def forward(self, gs):
        g, lg, _lat = gs                                                   # :294
        z = self.angle_embedding(lg.edata["h"])                            # :298
        x = self.atom_embedding(g.ndata["atom_features"])                  # :307-309
        y = self.edge_embedding(torch.norm(g.edata["r"], dim=1))           # :313-314
        x, y = self.conv_stack(g, lg, x, y, z)
        out = self.fc(avg_pool(g, x))                                      # :325,:341
        if self.link_name == "log":
            out = torch.exp(out)
        elif self.link_name == "logit":
            out = torch.sigmoid(out)
        if self.classification:
            out = self.softmax(out)
        return torch.squeeze(out)                                          # :349


def cutoff_envelope(r, inner_cutoff=4, exponent=3):
    """alignn/models/utils.py:58-86: polynomial envelope of x = r / inner_cutoff, zero beyond the cutoff."""
    ratio = r / inner_cutoff
    c1 = -(exponent + 1) * (exponent + 2) / 2
    c2 = exponent * (exponent + 2)
    c3 = -exponent * (exponent + 1) / 2
    env = 1 + c1 * ratio ** exponent + c2 * ratio ** (exponent + 1) + c3 * ratio ** (exponent + 2)
    return torch.where(r <= inner_cutoff, env, torch.zeros_like(r))


def energy_and_forces(model: ALIGNN, g: OGraph, lg: OGraph, energy_mult_natoms=True, use_penalty=True,
                      penalty_factor=0.1, penalty_threshold=1.0, use_cutoff_function=False, multiply_cutoff=False,
                      inner_cutoff=3.0, exponent=5, create_graph=False):
    """Energy + per-atom forces the ALIGNN-FF way (alignn_atomwise.py:404-467,495-510,526-563).

    r requires grad; cosines recomputed from r inside the autograd graph (lg_on_fly);
    optional cutoff envelope on the bond lengths (:434-451; without `multiply_cutoff` the envelope REPLACES the bond
    length, also for the penalty below);
    en = fc(avgpool(x)) [* natoms] + sum of short-bond penalties (:495-510; with energy_mult_natoms=False the reference
    adds the penalty in place to `out` itself, so it also appears in result["out"]);
    pair_forces = -dE/dr (grad_multiplier=-1); forces = sum_in-edges pair_forces - sum_out-edges pair_forces.
    """
    r = g.edata["r"].detach().clone().requires_grad_(True)
    h = bond_cosines(r, lg.src, lg.dst)
    z = model.angle_embedding(h)
    x = model.atom_embedding(g.ndata["atom_features"])
    bondlength = torch.norm(r, dim=1)
    if use_cutoff_function:
        env = cutoff_envelope(bondlength, inner_cutoff, exponent)
        if multiply_cutoff:
            y = model.edge_embedding(bondlength) * env.unsqueeze(1)
        else:
            bondlength = env
            y = model.edge_embedding(bondlength)
    else:
        y = model.edge_embedding(bondlength)
    x, y = model.conv_stack(g, lg, x, y, z)
    out = torch.squeeze(model.fc(avg_pool(g, x)))
    en = out * g.bnn.to(out.dtype) if energy_mult_natoms else out
    if use_penalty:
        pen = torch.where(bondlength < penalty_threshold, penalty_factor * (penalty_threshold - bondlength),
                          torch.zeros_like(bondlength))
        en = en + pen.sum()
        if not energy_mult_natoms:
            out = en                                  # the in-place `en_out += total_penalty` on the alias of `out`
    (dr,) = torch.autograd.grad(en.sum(), r, create_graph=create_graph)   # create_graph=True: force training (:530-539)
    pair_forces = -dr
    zeros = torch.zeros(g.n, 3, dtype=r.dtype)
    f_ji = zeros.index_add(0, g.dst, pair_forces)     # copy_e/sum on g           (:547-550)
    f_ij = zeros.index_add(0, g.src, pair_forces)     # copy_e/sum on reverse(g)  (:555-562)
    # result['out'] is the un-multiplied per-graph output (alignn_atomwise.py:653); en_out drives forces
    if create_graph:
        return out, f_ji - f_ij, pair_forces
    return out.detach(), (f_ji - f_ij).detach(), pair_forces.detach()


def virial_stress(g: OGraph, pair_forces, V, stress_multiplier=1.0):
    """Batched virial stress, alignn_atomwise.py:610-635: for crystal b, -(160.21766208 * r_b^T @ F_b / V[first atom
    of b]) -- the loop adds `num_nodes = 0` to `count_node` before indexing V, i.e. it reads the first atom's volume."""
    r = g.edata["r"]
    out, ce, cn = [], 0, 0
    for b in range(len(g.bne)):
        ne = int(g.bne[b])
        out.append(-1 * (160.21766208 * torch.matmul(r[ce:ce + ne].T, pair_forces[ce:ce + ne]) / V[cn]))
        ce += ne
        cn += int(g.bnn[b])
    return stress_multiplier * torch.stack(out)


def radius_graph(lattice_mat, cart_coords, cutoff=5.0, bond_tol=0.5, atol=1e-5, cutoff_extra=0.5):
    """Periodic radius graph, alignn/graphs.py:267-364, on torch tensors (cartesian_prod + cdist + where).

    Returns (u, v, r, cell_images) like the reference; retries with a larger cutoff until the highest-numbered
    atom has a bond (the reference's `g.num_nodes() == len(atoms.elements)` test, :347-350)."""
    import math
    X = torch.as_tensor(cart_coords, dtype=torch.float64)
    lat = torch.as_tensor(lattice_mat, dtype=torch.float64)
    frac = X @ torch.linalg.inv(lat)
    n = X.shape[0]
    while True:
        recp = 2 * math.pi * torch.linalg.inv(lat).T                       # :291
        recp_len = torch.sqrt(torch.sum(recp ** 2, dim=1))
        maxr = torch.ceil((cutoff + bond_tol) * recp_len / (2 * math.pi))  # :295
        nmin = torch.floor(torch.min(frac, dim=0)[0]) - maxr
        nmax = torch.ceil(torch.max(frac, dim=0)[0]) + maxr
        ranges = [torch.arange(float(a), float(b), dtype=torch.float64) for a, b in zip(nmin, nmax)]
        cells = torch.cartesian_prod(*ranges)                              # :304
        X_dst = ((cells @ lat)[:, None, :] + X).reshape(-1, 3)             # :308-311
        dist = torch.cdist(X, X_dst, compute_mode="donot_use_mm_for_euclid_dist")
        mask = (dist <= cutoff) & ~torch.isclose(dist, torch.zeros(1, dtype=torch.float64), atol=atol)   # :318-325
        u, v = torch.where(mask)
        images = cells[v // n]
        r = (X_dst[v] - X[u]).float()
        v = v % n
        if u.numel() and int(max(u.max(), v.max())) + 1 == n:
            return u, v, r, images
        cutoff += cutoff_extra


def knn_graph(lattice_mat, cart_coords, max_neighbors=12, cutoff=8.0):
    """k-NN crystal graph, alignn/graphs.py:155-264 with use_canonize=True, in plain Python (small cells only).

    `atoms.get_all_neighbors(r)` (jarvis-tools, absent) is restated as a brute-force search over periodic images.
    Ties are broken by (distance, v, image) and the images of one canonical pair are emitted in sorted order (the
    reference's order there is Python-set iteration order)."""
    import itertools
    from collections import OrderedDict
    lat = np.asarray(lattice_mat, dtype=np.float64)
    X = np.asarray(cart_coords, dtype=np.float64)
    n = X.shape[0]
    abc = np.linalg.norm(lat, axis=1)
    frac = X @ np.linalg.inv(lat)
    while True:
        reach = [int(np.ceil(cutoff / h)) + 1 for h in (abs(np.linalg.det(lat)) / np.array(
            [np.linalg.norm(np.cross(lat[1], lat[2])), np.linalg.norm(np.cross(lat[2], lat[0])),
             np.linalg.norm(np.cross(lat[0], lat[1]))]))]
        nbrs = [[] for _ in range(n)]
        for im in itertools.product(*[range(-m, m + 1) for m in reach]):
            shift = np.asarray(im, dtype=np.float64) @ lat
            for i in range(n):
                d = np.sqrt((((X + shift) - X[i]) ** 2).sum(1))
                for j in np.nonzero((d <= cutoff) & (d > 1e-8))[0]:
                    nbrs[i].append((float(d[j]), int(j), tuple(int(t) for t in im)))
        if min(len(l) for l in nbrs) >= max_neighbors:                     # :166-186
            break
        cutoff = float(abc.max()) if cutoff < abc.max() else 2 * cutoff
    edges = OrderedDict()
    for i, lst in enumerate(nbrs):
        lst = sorted(lst)                                                  # :202 (distance, then v, image)
        max_dist = lst[max_neighbors - 1][0]                               # :208
        for dist, j, im in lst:
            if dist > max_dist:                                            # :212-214 keep the whole shell
                continue
            a, b, img = (i, j, im) if j >= i else (j, i, tuple(-t for t in im))   # canonize_edge :127-152
            edges.setdefault((a, b), set()).add(img)
    u, v, r, images = [], [], [], []
    for (a, b), ims in edges.items():                                      # :240-257
        for im in sorted(ims):
            d = (frac[b] + np.asarray(im, dtype=np.float64) - frac[a]) @ lat
            for uu, vv, dd in ((a, b, d), (b, a, -d)):
                u.append(uu)
                v.append(vv)
                r.append(dd)
                images.append(im)
    return (np.asarray(u), np.asarray(v), np.asarray(r, dtype=np.float32), np.asarray(images).reshape(-1, 3))
