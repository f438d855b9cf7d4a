"""pipe vs plain chain kernel over a whole search (simulate path: graph + PDL): first diverging simulation."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, golden_io
from stochastic_muzero_b200 import ModelShape, SearchEngine
SEARCH = dict(pb_c_base=19652, pb_c_init=1.25, discount=0.997, root_dirichlet_alpha=0.25,
              root_exploration_fraction=0.25, num_simulations=50, maxium_action_sample=2, number_of_player=1, custom_loop=None)
z = golden_io.load_net_case("ckpt450")
B = 300
obs = torch.randn(B, 4, generator=torch.Generator().manual_seed(1)).cuda()
out = {}
for pipe in (0, 1):
    if pipe: os.environ.pop("SMZ_NO_PIPE", None)
    else: os.environ["SMZ_NO_PIPE"] = "1"
    eng = SearchEngine(SEARCH, 2, 2, max_trees=B, model_shape=ModelShape(4, 2, 2, 61, 126, 4), net="bf16", seed=99, record=True)
    eng.set_weights(z["weights"])
    eng.root(obs=obs, train=True)
    eng.simulate(50)
    torch.cuda.synchronize()
    try:
        eng.stats()
    except Exception as e:
        print("pipe", pipe, "stats:", e)
    out[pipe] = {k: v.cpu().numpy() for k, v in eng.read_record().items()}
    eng.close()
a, b = out[0], out[1]
for k in ("sim_policy", "sim_value", "sim_reward"):
    x, y = a[k], b[k]
    bad = ~np.isclose(x, y, equal_nan=False)
    if bad.ndim == 3: bad = bad.any(2)
    sims = np.flatnonzero(bad.any(0))
    print(k, "first bad sim:", sims[:5], "rows at first:", np.flatnonzero(bad[:, sims[0]])[:20] if len(sims) else None,
          "n rows", int(bad[:, sims[0]].sum()) if len(sims) else 0)
    if len(sims):
        s = sims[0]; r = np.flatnonzero(bad[:, s])[0]
        print("   plain", x[r, s], "pipe", y[r, s])
