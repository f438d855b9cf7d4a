"""Vision family: simulation step with the MLP heads on the tensor cores against the all-CUDA-core kernel (SMZ_VISION_CC=1)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, golden_io
from stochastic_muzero_b200 import SearchEngine, VisionShape
z = golden_io.load_vision_case("a4")
A, S, H, L = [int(v) for v in z["dims"]]
B, N, seed = 300, 20, 5
search = dict(pb_c_base=19652, pb_c_init=1.25, discount=0.997, root_dirichlet_alpha=0.25, root_exploration_fraction=0.25,
              num_simulations=N, maxium_action_sample=2, number_of_player=1, custom_loop=None)
obs = torch.rand(B, 3, 98, 98, generator=torch.Generator().manual_seed(8)).reshape(B, -1)
out = {}
for mode in ("tc", "cc"):
    os.environ.pop("SMZ_VISION_CC", None)
    if mode == "cc": os.environ["SMZ_VISION_CC"] = "1"
    eng = SearchEngine(search, A, A, max_trees=B, model_shape=VisionShape(A, S, H, L), net="vision", rng="philox", seed=seed, record=True)
    eng.set_weights(z["weights"])
    eng.root(obs=obs, train=True); eng.simulate(N); print(mode, eng.stats())
    rec = {k: v.cpu().numpy() for k, v in eng.read_record().items()}
    out[mode] = (rec, eng.read_roots()["visits"].cpu().numpy(), eng.read_hidden(N).cpu().numpy())
    eng.close()
a, b = out["tc"][0], out["cc"][0]
samev = (out["tc"][1] == out["cc"][1]).all(1); sameb = (a["sim_branch"] == b["sim_branch"]).all(1)
print("same visits", samev.mean(), "same branch", sameb.mean())
# first simulation is identical for all trees (same root): compare sim 0 outputs for every tree
for k in ("sim_policy", "sim_value", "sim_reward"):
    d = np.abs(a[k][:, 0] - b[k][:, 0]); print(k, "sim0 max diff", d.max(), "at", np.unravel_index(d.argmax(), d.shape))
same = samev & sameb
for k in ("sim_policy", "sim_value", "sim_reward"):
    d = np.abs(a[k][same] - b[k][same]); print(k, "max diff on same trees", d.max(), "scale", np.abs(b[k]).max())
bad = np.flatnonzero(~same)[:10]; print("bad trees", bad)
for t in bad[:3]:
    fs = np.flatnonzero((a["sim_branch"][t] != b["sim_branch"][t]) | (np.abs(a["sim_value"][t]-b["sim_value"][t])>1e-4))
    print(t, "first differing sim", fs[:3], a["sim_value"][t][:4], b["sim_value"][t][:4])
