import os, sys
os.environ["SMZ_BF16_TIMELINE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, golden_io
from stochastic_muzero_b200 import ModelShape, SearchEngine
SEARCH = dict(pb_c_base=19652, pb_c_init=1.25, discount=0.997, root_dirichlet_alpha=0.25,
              root_exploration_fraction=0.25, num_simulations=50, maxium_action_sample=2, number_of_player=1, custom_loop=None)
z = golden_io.load_net_case("mlp450_seed0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
eng = SearchEngine(SEARCH, 2, 2, max_trees=B, model_shape=ModelShape(4, 2, 2, 61, 126, 4), net="bf16", seed=7)
eng.set_weights(z["weights"])
obs = torch.randn(B, 4).cuda()
for _ in range(3):
    eng.root(obs=obs, train=True); eng.simulate(50)
torch.cuda.synchronize()
eng.close()
