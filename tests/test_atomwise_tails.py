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


# ---- force TRAINING: double backward through the torch-operator composition (conv.second_order) --------------------
class _TorchOpsLayer(torch.nn.Module):
    """The product layer run through `conv._torch_ops_forward` (what `second_order` selects on the GPU), on the CPU."""

    def __init__(self, layer, kind):
        super().__init__()
        self.layer, self.kind = layer, kind

    def forward(self, g, *args, **kw):
        from alignn_b200 import conv
        if self.kind == "gcn":
            x, y = args
            return conv._torch_ops_forward(self.layer, g.index, x, y, kw.get("_need_edge_out", True))
        lg, x, y, z = args
        x, m = conv._torch_ops_forward(self.layer.node_update, g.index, x, y, True)
        y, z = conv._torch_ops_forward(self.layer.edge_update, lg.index, m, z, kw.get("_need_z_out", True))
        return x, y, z


def test_force_training_gradients_match_oracle_double_backward(cpu_pool):
    """d(force loss + energy loss)/d(parameters) through create_graph=True (alignn_atomwise.py:530-539) agrees with the
    oracle's double backward: the force term really trains (it silently did not)."""
    g, lg, lat = _batch()
    m = A.ALIGNNAtomWise(A.ALIGNNAtomWiseConfig(name="alignn_atomwise", alignn_layers=2, gcn_layers=2, hidden_features=32,
                                                embedding_features=32, atom_input_features=92)).double().train()
    GI.fill_state_dict(m, 77)
    orc = O.ALIGNN(norm="layernorm", alignn_layers=2, gcn_layers=2, hidden_features=32, embedding_features=32).double().train()
    orc.load_state_dict(m.state_dict())
    m.alignn_layers = torch.nn.ModuleList([_TorchOpsLayer(l, "alignn") for l in m.alignn_layers])
    m.gcn_layers = torch.nn.ModuleList([_TorchOpsLayer(l, "gcn") for l in m.gcn_layers])
    res = m((g, lg, lat))
    assert res["grad"].requires_grad                       # the forces stay in the autograd graph in training
    tgt_f = GI.features(12, g.num_nodes(), 3).double()
    loss = (res["grad"] - tgt_f).abs().mean() + res["out"].abs().mean()
    loss.backward()
    out, forces, _ = O.energy_and_forces(orc, to_oracle(g), to_oracle(lg), create_graph=True)
    ref = (forces - tgt_f).abs().mean() + out.abs().mean()
    ref.backward()
    got = {k.replace(".layer.", ".", 1) if False else k: v for k, v in m.named_parameters()}
    n_checked = 0
    for name, p_ref in orc.named_parameters():
        name_m = name
        for pre in ("alignn_layers.", "gcn_layers."):
            if name.startswith(pre):
                i, rest = name[len(pre):].split(".", 1)
                name_m = f"{pre}{i}.layer.{rest}"
        p = dict(m.named_parameters())[name_m]
        if p_ref.grad is None:
            assert p.grad is None or p.grad.abs().max() == 0
            continue
        assert p.grad is not None, name
        scale = max(p_ref.grad.abs().max().item(), 1e-12)
        assert (p.grad - p_ref.grad).abs().max().item() <= 1e-7 * scale + 1e-14, name
        n_checked += 1
    assert n_checked > 30
    # the force term contributes: a conv weight's gradient differs from the energy-only gradient
    assert dict(m.named_parameters())["gcn_layers.0.layer.src_gate.weight"].grad.abs().max() > 0


def test_property_only_config_skips_the_force_pass():
    """alignn_atomwise.py:267-268: gradwise_weight == 0 switches calculate_gradient off (works under no_grad)."""
    cfg = A.ALIGNNAtomWiseConfig(name="alignn_atomwise", gradwise_weight=0.0)
    m = A.ALIGNNAtomWise(cfg)
    assert m.config.calculate_gradient is False
