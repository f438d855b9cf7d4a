"""How often does the bf16 tensor-core network step change the SEARCH RESULT relative to the fp32 step?
Same seeds, same observations; trained 450 checkpoint and random-init weights."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, golden_io
from stochastic_muzero_b200 import ModelShape, SearchEngine
SEARCH = dict(pb_c_base=19652, pb_c_init=1.25, discount=0.997, root_dirichlet_alpha=0.25,
              root_exploration_fraction=0.25, num_simulations=50, maxium_action_sample=2, number_of_player=1, custom_loop=None)
for name in ("mlp450_seed0", "ckpt450"):
    z = golden_io.load_net_case(name)
    B = 4096
    obs = (torch.randn(B, 4, generator=torch.Generator().manual_seed(0)) * 0.1).cuda()
    res = {}
    for net in ("fp32", "bf16"):
        eng = SearchEngine(SEARCH, 2, 2, max_trees=B, model_shape=ModelShape(4, 2, 2, 61, 126, 4), net=net, seed=11)
        eng.set_weights(z["weights"])
        eng.root(obs=obs, train=True); eng.simulate(50)
        r = eng.read_roots()
        res[net] = (r["visits"].cpu().numpy(), r["root_values"].cpu().numpy())
        eng.close()
    v32, v16 = res["fp32"][0], res["bf16"][0]
    same = (v32 == v16).all(1).mean()
    argsame = (v32.argmax(1) == v16.argmax(1)).mean()
    dv = np.abs(v32 - v16).max(1)
    rel = np.abs(res["fp32"][1] - res["bf16"][1]) / np.maximum(np.abs(res["fp32"][1]), 1e-6)
    print(f"{name}: identical visit vectors {same:.4f}; same most-visited action {argsame:.4f}; max |dvisit| mean {dv.mean():.2f} "
          f"p99 {np.percentile(dv, 99):.0f}; root value rel diff median {np.median(rel):.2e} p99 {np.percentile(rel, 99):.2e}")
