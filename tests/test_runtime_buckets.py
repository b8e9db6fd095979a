"""Shape-bucketed CUDA-graph replay (alignn_b200/runtime.py): the padding crystal has exactly the requested numbers of
atoms / bonds / bond pairs (CPU), and on the GPU a LayerNorm model replayed from a bucket's graph gives bit-identical
predictions for the real crystals of differently sized batches."""
import numpy as np
import pytest
import torch

from alignn_b200 import runtime, synthetic
from oracle import golden_inputs as GI


@pytest.mark.parametrize("dN,dE,dT", [(8, 0, 0), (9, 5, 0), (30, 40, 391), (64, 300, 20000), (8, 3, 2), (12, 7, 10)])
def test_padding_crystal_has_exact_counts(dN, dE, dT):
    need_n, need_e = runtime.padding_needs(dE, dT)
    if dE < need_e:
        with pytest.raises(ValueError):
            runtime.make_padding_crystal(dN, dE, dT, 4)
        dE = need_e
    g = runtime.make_padding_crystal(dN, dE, dT, 4)
    lg = g.line_graph()
    assert (g.num_nodes(), g.num_edges(), lg.num_edges()) == (dN, dE, dT)


def test_pad_batch_keeps_the_real_crystals_in_front():
    g, lg, lat, _ = synthetic.make_batch(batch_size=3, atoms=9, k=12, seed=3, vary_atoms=True)
    N, E, T = g.num_nodes(), g.num_edges(), lg.num_edges()
    pb = runtime.pad_batch(g, lg, lat, N + 20, E + 100, T + 1234)
    assert (pb.g.num_nodes(), pb.g.num_edges(), pb.lg.num_edges()) == (N + 20, E + 100, T + 1234)
    assert pb.num_real == 3 and pb.g.batch_size == 4
    s, d = g.edges()
    ps, pd = pb.g.edges()
    assert torch.equal(ps[:E], s) and torch.equal(pd[:E], d) and torch.equal(pb.g.edata["r"][:E], g.edata["r"])
    ls, ld = lg.edges()
    pls, pld = pb.lg.edges()
    assert torch.equal(pls[:T], ls) and torch.equal(pld[:T], ld)
    assert torch.equal(pb.lg.edata["h"][:T], lg.edata["h"])
    assert torch.isfinite(pb.lg.edata["h"]).all()


@pytest.mark.gpu
def test_bucketed_graph_replay_bit_identical_to_eager_for_layernorm_model():
    from alignn_b200 import alignn_atomwise as AW
    from alignn_b200.alignn import ALIGNN, ALIGNNConfig

    class Model(ALIGNN):
        _mlp, _alignn_conv, _gcn_conv = AW.MLPLayer, AW.ALIGNNConv, AW.EdgeGatedGraphConv
    dev = torch.device("cuda:0")
    m = Model(ALIGNNConfig(name="alignn", alignn_layers=2, gcn_layers=2, hidden_features=64, embedding_features=32))
    GI.fill_state_dict(m, 8)
    m.to(dev).eval()

    def fn(g, lg, lat):
        with torch.no_grad():
            return m((g, lg, lat))
    runner = runtime.BucketedForward(fn, dev, n_edges=runtime.geometric_buckets(64, 400), e_edges=runtime.geometric_buckets(512, 6000),
                                     t_edges=runtime.geometric_buckets(4096, 90000))
    seen = set()
    for seed in (1, 2, 3, 4, 5):
        g, lg, lat, _ = synthetic.make_batch(batch_size=4, atoms=9, k=12, seed=seed, vary_atoms=True)
        out, n_real = runner(g, lg, lat)
        seen.add(runner.bucket(g, lg))
        with torch.no_grad():
            ref = m((g.to(dev), lg.to(dev), lat.to(dev)))
        assert n_real == 4 and torch.equal(out[:n_real], ref)
    assert len(runner._graphs) == len(seen) <= 4          # several batches share a bucket (and its graph)
