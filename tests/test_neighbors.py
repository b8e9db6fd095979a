"""CPU: the host-side periodic radius-graph builder (feeds BASELINE config 4) against the oracle restatement of
alignn/graphs.py:267-364, bond for bond and in the same order."""
import numpy as np
import pytest
import torch

from alignn_b200 import neighbors
from oracle import alignn_oracle as O


def _random_cell(seed, n, skew=0.3):
    rng = np.random.default_rng(seed)
    lat = np.eye(3) * (4.0 + 2.0 * rng.random(3)) + skew * rng.normal(size=(3, 3))
    frac = rng.random((n, 3))
    return lat, frac @ lat


@pytest.mark.parametrize("seed,n,cutoff", [(1, 1, 5.0), (2, 2, 4.0), (3, 7, 4.5), (4, 12, 3.5)])
def test_radius_graph_matches_reference_algorithm(seed, n, cutoff):
    lat, X = _random_cell(seed, n)
    u, v, r, im = neighbors.radius_graph(lat, X, cutoff=cutoff)
    uo, vo, ro, imo = O.radius_graph(lat, X, cutoff=cutoff)
    assert np.array_equal(u, uo.numpy()) and np.array_equal(v, vo.numpy())          # same bonds, same order
    assert np.array_equal(im, imo.numpy())
    np.testing.assert_allclose(r, ro.numpy(), rtol=0, atol=1e-5)
    d = np.linalg.norm(r, axis=1)
    assert d.min() > 1e-5 and d.max() <= max(cutoff, d.max())                        # no self distance
    # every bond has its reverse (same multiset of displacement lengths per unordered pair)
    fwd = sorted(zip(u.tolist(), v.tolist(), np.round(d, 4).tolist()))
    rev = sorted(zip(v.tolist(), u.tolist(), np.round(d, 4).tolist()))
    assert fwd == rev


def test_diamond_supercell_is_16_coordinated_within_4_angstrom():
    """Diamond Si: 4 first + 12 second neighbours inside 4 A (E = 16 N, T = sum deg^2)."""
    lat, X = neighbors.diamond_supercell(reps=2)
    assert X.shape == (64, 3)
    g, lg = neighbors.crystal_graph(lat, X, torch.zeros(64, 92), cutoff=4.0)
    assert g.num_edges() == 64 * 16
    indeg = torch.bincount(g.edges()[1].long(), minlength=64)
    assert int(indeg.min()) == 16 and int(indeg.max()) == 16
    assert lg.num_edges() == 64 * 16 * 16                     # no self loops in g -> every (i, j) pair kept
    assert lg.index.dst_sorted
    h = lg.edata["h"]
    assert float(h.min()) >= -1.0 and float(h.max()) <= 1.0
    # tetrahedral angle between first-neighbour bonds: cos = -1/3 must occur
    assert torch.isclose(h, torch.tensor(-1.0 / 3.0), atol=1e-5).any()


@pytest.mark.parametrize("seed,n,k", [(1, 1, 12), (2, 2, 12), (3, 5, 8), (4, 9, 12)])
def test_knn_graph_matches_reference_algorithm(seed, n, k):
    """k-NN strategy (the reference default, graphs.py:155-264): same bonds in the same order as the plain-Python
    restatement; both directions adjacent; every atom keeps whole shells (degree >= k)."""
    lat, X = _random_cell(seed, n, skew=0.2)
    u, v, r, im = neighbors.knn_graph(lat, X, max_neighbors=k, cutoff=4.0)
    uo, vo, ro, imo = O.knn_graph(lat, X, max_neighbors=k, cutoff=4.0)
    assert np.array_equal(u, uo) and np.array_equal(v, vo) and np.array_equal(im, imo)
    np.testing.assert_allclose(r, ro, atol=1e-5)
    assert np.array_equal(u[0::2], v[1::2]) and np.array_equal(v[0::2], u[1::2])     # reverse bond adjacent
    np.testing.assert_allclose(r[0::2], -r[1::2], atol=1e-6)
    assert np.bincount(v, minlength=n).min() >= k


def test_knn_graph_on_diamond_gives_first_two_shells():
    lat, X = neighbors.diamond_supercell(reps=2)
    u, v, r, im = neighbors.knn_graph(lat, X, max_neighbors=12, cutoff=8.0)
    deg = np.bincount(v, minlength=64)
    # 4 + 12: the 12th neighbour lies in the second shell.  Like the reference (`if dist > max_dist`, graphs.py:212)
    # the shell test compares doubles exactly, so second-shell members whose distance rounds 1 ulp above the 12th
    # one are dropped unless the partner atom kept the bond: degrees land between 12 and 16.
    assert deg.min() >= 12 and deg.max() <= 16
    d = np.linalg.norm(r, axis=1)
    assert np.allclose(np.unique(np.round(d, 3)), [2.352, 3.840])


@pytest.mark.parametrize("seed", [5, 6])
def test_radius_graph_is_invariant_under_rigid_translation_and_atom_relabelling(seed):
    """Size-independent properties of the periodic neighbour search: the multiset of bond lengths does not change when
    the crystal is translated rigidly, and relabelling atoms only relabels bonds."""
    lat, X = _random_cell(seed, 6)
    _, _, r0, _ = neighbors.radius_graph(lat, X, cutoff=4.0)
    base = np.sort(np.round(np.linalg.norm(r0, axis=1), 5))
    shift = np.array([0.37, -1.2, 2.9])
    _, _, r1, _ = neighbors.radius_graph(lat, X + shift, cutoff=4.0)
    assert np.array_equal(np.sort(np.round(np.linalg.norm(r1, axis=1), 5)), base)
    perm = np.random.default_rng(seed).permutation(6)
    u2, v2, r2, _ = neighbors.radius_graph(lat, X[perm], cutoff=4.0)
    assert np.array_equal(np.sort(np.round(np.linalg.norm(r2, axis=1), 5)), base)
    # bond (u, v) of the relabelled crystal joins the original atoms (perm[u], perm[v]): displacement = image offset apart
    d = (X[perm][v2] - X[perm][u2]) - r2
    frac = d @ np.linalg.inv(lat)
    assert np.allclose(frac, np.round(frac), atol=1e-5)


def test_crystal_graph_with_the_k_nearest_strategy_builds_graph_and_line_graph():
    lat, X = neighbors.diamond_supercell(reps=2, jitter=0.02, seed=1)
    g, lg = neighbors.crystal_graph(lat, X, torch.zeros(64, 92), cutoff=8.0, neighbor_strategy="k-nearest", max_neighbors=12)
    s, t = (a.long() for a in g.edges())
    assert torch.equal(s[0::2], t[1::2]) and torch.equal(t[0::2], s[1::2])       # (u, v) then (v, u), graphs.py:253-257
    assert int(torch.bincount(t, minlength=64).min()) >= 12
    assert lg.num_nodes() == g.num_edges() and lg.index.dst_sorted
    with pytest.raises(ValueError):
        neighbors.crystal_graph(lat, X, torch.zeros(64, 92), neighbor_strategy="voronoi")
