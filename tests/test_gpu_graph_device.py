"""GPU (-m gpu): the device-side structure builders and ALIGNN-FF reductions of the shipped library
(csrc/graph_device.cu) against the native host builders (bit-identical integers) and fp64 restatements."""
import math

import numpy as np
import pytest
import torch

from alignn_b200 import neighbors, ops, synthetic
from alignn_b200.graph import EdgeIndex, Graph
from oracle import alignn_oracle as O
from oracle import golden_inputs as GI
from tests.helpers import to_oracle

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _test_graphs():
    g, lg, _, _ = synthetic.make_batch(batch_size=3, atoms=6, k=12, seed=31, vary_atoms=True)
    s, t = (a.numpy() for a in g.edges())
    perm = np.random.default_rng(2).permutation(s.size)
    shuffled = Graph(s[perm], t[perm], g.num_nodes(), g.batch_num_nodes(), g.batch_num_edges())
    loops = Graph(np.array([0, 1, 1, 2, 2, 2, 0]), np.array([1, 1, 0, 2, 0, 2, 0]), 4)      # self-loops, isolated node 3
    empty = Graph(np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64), 3)
    return [g, lg, shuffled, loops, empty]


def test_device_csr_and_line_graph_bit_identical_to_host_builders():
    for gr in _test_graphs():
        ix = gr.index                                   # native host builder
        dix = EdgeIndex.build_device(ix.src.to(DEV), ix.dst.to(DEV), gr.num_nodes())
        for k in ("in_ptr", "in_eid", "out_ptr", "out_eid"):
            assert torch.equal(getattr(dix, k).cpu(), getattr(ix, k)), k
        assert dix.dst_sorted == ix.dst_sorted and dix.max_in_deg == ix.max_in_deg
        ref = gr.line_graph()                           # host
        dlg = gr.to(DEV).line_graph()                   # device (Graph.line_graph dispatches on the graph's device)
        assert dlg.device.type == "cuda"
        rs, rt = ref.edges()
        ds, dt = dlg.edges()
        assert torch.equal(ds.cpu(), rs) and torch.equal(dt.cpu(), rt)
        for k in ("in_ptr", "in_eid", "out_ptr", "out_eid"):
            assert torch.equal(getattr(dlg.index, k).cpu(), getattr(ref.index, k)), k
        assert torch.equal(dlg.batch_num_edges(), ref.batch_num_edges())
        assert dlg.index.dst_sorted == ref.index.dst_sorted


def test_device_force_scatter_and_virial_match_fp64():
    g, _, _, _ = synthetic.make_batch(batch_size=3, atoms=7, k=12, seed=37, vary_atoms=True)
    E = g.num_edges()
    pf = GI.features(8, E, 3)
    gd = g.to(DEV)
    f = ops.pair_force_scatter(pf.to(DEV), gd.index).cpu().double()
    s, t = (a.long() for a in g.edges())
    zeros = torch.zeros(g.num_nodes(), 3, dtype=torch.float64)
    ref = zeros.index_add(0, t, pf.double()) - zeros.index_add(0, s, pf.double())
    assert (f - ref).abs().max() <= 1e-5 * ref.abs().max()
    vols = GI.cell_volumes(g.batch_num_nodes())
    st = ops.virial_stress(g.edata["r"].to(DEV), pf.to(DEV), gd.edge_graph_offsets64(), gd.node_graph_offsets().long(),
                           vols.to(DEV), 10.0).cpu().double()
    og = to_oracle(g, torch.float64)
    ref = O.virial_stress(og, pf.double(), vols.double(), stress_multiplier=10.0)
    assert (st - ref).abs().max() <= 1e-5 * ref.abs().max()


@pytest.mark.parametrize("reps,jitter", [(1, 0.0), (2, 0.05), (3, 0.1)])
def test_device_radius_graph_bit_identical_to_host_scan(reps, jitter):
    """Same bonds, same order, same float32 displacement vectors as the native host scan (graphs.py:267-364)."""
    lat, X = neighbors.diamond_supercell(reps=reps, jitter=jitter, seed=4)
    u, v, r, cells_of_bond = neighbors.radius_graph(lat, X, cutoff=4.0)
    du, dv, dr, dc, cells = neighbors.radius_graph_device(lat, X, cutoff=4.0, device=DEV)
    assert np.array_equal(du.cpu().numpy(), u) and np.array_equal(dv.cpu().numpy(), v)
    assert np.array_equal(cells[dc.cpu().numpy()], cells_of_bond)
    assert np.array_equal(dr.cpu().numpy(), r)


def test_crystal_graph_device_equals_host_pipeline():
    lat, X = neighbors.diamond_supercell(reps=2, jitter=0.05, seed=9)
    feats = GI.features(3, X.shape[0], 92)
    g, lg = neighbors.crystal_graph(lat, X, feats, cutoff=4.0)
    dg, dlg = neighbors.crystal_graph_device(lat, X, feats, cutoff=4.0, device=DEV)
    for a, b in ((g, dg), (lg, dlg)):
        s, t = a.edges()
        ds, dt = b.edges()
        assert torch.equal(ds.cpu(), s) and torch.equal(dt.cpu(), t)
    assert torch.equal(dg.edata["r"].cpu(), g.edata["r"])
    assert torch.allclose(dlg.edata["h"].cpu(), lg.edata["h"], rtol=0, atol=1e-6)
