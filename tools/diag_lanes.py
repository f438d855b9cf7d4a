import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, golden_io
from stochastic_muzero_b200 import ModelShape, SearchEngine
SEARCH = dict(pb_c_base=19652, pb_c_init=1.25, discount=0.997, root_dirichlet_alpha=0.25,
              root_exploration_fraction=0.25, num_simulations=50, maxium_action_sample=2, number_of_player=1, custom_loop=None)
z = golden_io.load_net_case("mlp450_seed0")
for B in (4096, 16384, 32768):
    obs = torch.randn(B, 4).cuda()
    for lanes in (2, 4, 8, 16, 32):
        eng = SearchEngine(SEARCH, 2, 2, max_trees=B, model_shape=ModelShape(4, 2, 2, 61, 126, 4), net="bf16", seed=7, lanes_per_tree=lanes)
        eng.set_weights(z["weights"])
        for _ in range(3):
            eng.root(obs=obs, train=True); eng.simulate(50)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(10):
            eng.set_seed(100 + i); eng.root(obs=obs, train=True); eng.simulate(50)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        print(f"B={B} lanes={lanes}: {ms:.3f} ms/search = {B*50/ms/1e3:.1f} M sims/s", flush=True)
        eng.close()
