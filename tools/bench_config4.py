"""BASELINE config 4: ALIGNN-FF energy + per-atom forces on a ~1000-atom periodic supercell, 1 GPU.
(Secondary config: reported in DESIGN.md / profiles, not the headline bench line.)"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from alignn_b200 import neighbors  # noqa: E402
from alignn_b200.alignn_atomwise import ALIGNNAtomWise, ALIGNNAtomWiseConfig  # noqa: E402

dev = torch.device("cuda:0")
t0 = time.perf_counter()
lat, X = neighbors.diamond_supercell(reps=5, jitter=0.03)
g, lg = neighbors.crystal_graph(lat, X, torch.rand(X.shape[0], 92), cutoff=4.0)
t_build = time.perf_counter() - t0
torch.manual_seed(0)
m = ALIGNNAtomWise(ALIGNNAtomWiseConfig(name="alignn_atomwise", atom_input_features=92, alignn_layers=4, gcn_layers=4,
                                        hidden_features=256)).to(dev).eval()
gd, lgd, latd = g.to(dev), lg.to(dev), torch.from_numpy(lat).float().unsqueeze(0).to(dev)


def run():
    return m((gd, lgd, latd))


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    res = run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(json.dumps({"config4_alignn_ff": {"atoms": g.num_nodes(), "bonds": g.num_edges(), "triplets": lg.num_edges(),
                                         "ms_per_energy_force_eval": ms, "evals_per_s": 1e3 / ms,
                                         "host_graph_build_s": t_build,
                                         "max_force": float(res["grad"].abs().max())}}))
