"""cfg3-shaped search for profiling the tree step (8192 trees, A 4, C 32, K 32): a few simulations, step by step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stochastic_muzero_b200 import ModelShape, SearchEngine
from stochastic_muzero_b200.weights import random_blob
shape = ModelShape(obs_dim=16, action_dim=4, chance_dim=32, state_dim=61, hidden_dim=126, num_hidden_layers=4)
search = dict(pb_c_base=19652, pb_c_init=1.25, discount=0.997, root_dirichlet_alpha=0.25, root_exploration_fraction=0.25,
              num_simulations=100, maxium_action_sample=32, number_of_player=1, custom_loop=None)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
eng = SearchEngine(search, 4, 32, max_trees=B, model_shape=shape, net="bf16", rng="philox", seed=3)
eng.set_weights(random_blob(shape, seed=0))
obs = (torch.randint(0, 12, (B, 16)).float() / 16.0).cuda()
eng.root(obs=obs, train=True)
eng.select(0)
for s in range(60):
    eng.net_step(s)
    eng.backup_select(s)
torch.cuda.synchronize()
print("done", eng.stats())
