"""Where the end-to-end call spends its time beyond the device-timed search (cfg2 shapes)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stochastic_muzero_b200 import ModelShape, Monte_carlo_tree_search, PackedModel
from stochastic_muzero_b200.weights import random_blob
SEARCH = dict(pb_c_base=19652, pb_c_init=1.25, discount=0.997, root_dirichlet_alpha=0.25, root_exploration_fraction=0.25,
              num_simulations=50, maxium_action_sample=2, number_of_player=1, custom_loop=None)
shape = ModelShape(obs_dim=4, action_dim=2, chance_dim=2, state_dim=61, hidden_dim=126, num_hidden_layers=4)
B = 4096
model = PackedModel(random_blob(shape, seed=0), shape)
mcts = Monte_carlo_tree_search(**SEARCH, net="bf16", seed=7, max_batch=B)
obs_host = torch.randn(B, 4).pin_memory()
obs_dev = obs_host.cuda()
def loop(fn, n=40):
    for _ in range(5): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return 1e3 * (time.perf_counter() - t0) / n
print("device obs, no read-back, no per-step sync : %.4f ms" % loop(lambda: mcts.run_batch(obs_dev, model, train=True, check=False)))
print("host obs,   no read-back, no per-step sync : %.4f ms" % loop(lambda: mcts.run_batch(obs_host, model, train=True, check=False)))
print("host obs,   check=True (one copy + sync)   : %.4f ms" % loop(lambda: mcts.run_batch(obs_host, model, train=True)))
print("host obs,   .host()                        : %.4f ms" % loop(lambda: mcts.run_batch(obs_host, model, train=True).host()))
t0 = time.perf_counter()
for _ in range(200): mcts._is_fusable(model)
print("python: _is_fusable %.1f us" % (1e6 * (time.perf_counter() - t0) / 200))
