"""CPU: everything of `alignn_b200.ALIGNNAtomWise.forward` AROUND the conv stack -- embeddings, cutoff envelope, penalty
(with the reference's aliasing), pair-force reduction, virial stress, result dict -- against the unmodified reference's
outputs (tests/golden/atomwise_*.npz).  The CUDA conv layers are substituted by the oracle's conv layers carrying the
same weights (test-only; the product itself has no CPU path), so a mistake in the torch code around the kernels shows
up here without a GPU."""
import numpy as np
import pytest
import torch

from alignn_b200 import alignn_atomwise as A
from alignn_b200 import synthetic
from oracle import alignn_oracle as O
from oracle import golden_inputs as GI
from tests.helpers import to_oracle


class _OracleConv(torch.nn.Module):
    """Oracle conv with the product layer's parameters, taking the product's Graph arguments."""

    def __init__(self, product_layer, kind):
        super().__init__()
        self.kind = kind
        d = product_layer.src_gate.in_features if kind == "gcn" else product_layer.node_update.src_gate.in_features
        self.inner = (O.EdgeGatedGraphConv(d, d, norm="layernorm") if kind == "gcn" else O.ALIGNNConv(d, d, norm="layernorm")).double()
        self.inner.load_state_dict({k: v.double() for k, v in product_layer.state_dict().items()})

    def forward(self, g, *args, **_unused):
        if self.kind == "gcn":
            x, y = args
            return self.inner(to_oracle(g), x, y)
        lg, x, y, z = args
        return self.inner(to_oracle(g), to_oracle(lg), x, y, z)


def _cpu_model(**cfg):
    m = A.ALIGNNAtomWise(A.ALIGNNAtomWiseConfig(name="alignn_atomwise", alignn_layers=2, gcn_layers=2, hidden_features=64,
                                                embedding_features=32, atom_input_features=92, **cfg))
    m = m.double().eval()
    GI.fill_state_dict(m, 400)                  # filled in fp64 like the fixture generator (oracle/make_golden.py)
    m.alignn_layers = torch.nn.ModuleList([_OracleConv(l, "alignn") for l in m.alignn_layers])
    m.gcn_layers = torch.nn.ModuleList([_OracleConv(l, "gcn") for l in m.gcn_layers])
    return m


@pytest.fixture()
def cpu_pool(monkeypatch):
    def seg_mean(x, off):
        off = off.long()
        cnt = off[1:] - off[:-1]
        gid = torch.repeat_interleave(torch.arange(cnt.numel()), cnt)
        return torch.zeros(cnt.numel(), x.shape[1], dtype=x.dtype).index_add(0, gid, x) / cnt.to(x.dtype).unsqueeze(1)
    monkeypatch.setattr(A.ops, "segment_mean", seg_mean)


def _batch(scale=1.0):
    g, lg, lat, _ = synthetic.make_batch(batch_size=2, atoms=8, k=12, seed=41, vary_atoms=True)
    for gr in (g, lg):
        for dct in (gr.ndata, gr.edata):
            for k in list(dct):
                if dct[k].is_floating_point():
                    dct[k] = dct[k].double()
    g.edata["r"] = g.edata["r"] * scale
    return g, lg, lat.double()


def test_energy_forces_and_stress_tails_match_reference(golden_dir, cpu_pool):
    gold = np.load(f"{golden_dir}/atomwise_small.npz")
    gold_s = np.load(f"{golden_dir}/atomwise_stress.npz")
    g, lg, lat = _batch()
    g.ndata["V"] = GI.cell_volumes(g.batch_num_nodes()).double()
    res = _cpu_model(stresswise_weight=0.1, stress_multiplier=10.0)((g, lg, lat))
    np.testing.assert_allclose(res["out"].detach().numpy(), gold["out"], rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(res["grad"].numpy(), gold["forces"], rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(res["pair_forces"].numpy(), gold["pair_forces"], rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(res["stresses"].numpy(), gold_s["stresses"], rtol=1e-7, atol=1e-8)
    assert set(res) >= {"out", "grad", "stresses", "atomwise_pred", "additional"}        # alignn_atomwise.py:653-657


@pytest.mark.parametrize("tag,cfg", [
    ("mult", dict(use_cutoff_function=True, multiply_cutoff=True, inner_cutoff=2.5, exponent=5)),
    ("repl", dict(use_cutoff_function=True, multiply_cutoff=False, inner_cutoff=2.5, exponent=3)),
    ("leak", dict(use_cutoff_function=False, penalty_threshold=1.2, energy_mult_natoms=False))])
def test_cutoff_and_penalty_tails_match_reference(golden_dir, cpu_pool, tag, cfg):
    gold = np.load(f"{golden_dir}/atomwise_cutoff.npz")
    g, lg, lat = _batch(scale=0.35)
    res = _cpu_model(**cfg)((g, lg, lat))
    np.testing.assert_allclose(res["out"].detach().numpy(), gold[tag + ".out"], rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(res["grad"].numpy(), gold[tag + ".forces"], rtol=1e-7, atol=1e-8)
