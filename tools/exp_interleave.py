"""Experiment: does splitting the 4096-tree batch into independent groups that run
concurrently on separate streams shorten the search?  Times (a) one engine of B trees for several B,
(b) G engines of 4096/G trees launched together on G streams.  CUDA events, 20 repetitions, bf16 step."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stochastic_muzero_b200 import ModelShape, SearchEngine  # noqa: E402
from stochastic_muzero_b200.weights import random_blob  # noqa: E402

SEARCH = dict(pb_c_base=19652, pb_c_init=1.25, discount=0.997, root_dirichlet_alpha=0.25,
              root_exploration_fraction=0.25, num_simulations=50, maxium_action_sample=2,
              number_of_player=1, custom_loop=None)
shape = ModelShape(obs_dim=4, action_dim=2, chance_dim=2, state_dim=61, hidden_dim=126, num_hidden_layers=4)
net = sys.argv[1] if len(sys.argv) > 1 else "bf16"
blob = random_blob(shape, seed=0)
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def make(B, off):
    e = SearchEngine(SEARCH, 2, 2, max_trees=B, model_shape=shape, net=net, rng="philox", seed=7, tree_id_offset=off)
    e.set_weights(blob)
    return e


def timed(engs, streams, obs, reps=20):
    ms = []
    for it in range(reps + 3):
        flush.zero_()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        main = torch.cuda.current_stream()
        for e, s, o in zip(engs, streams, obs):
            s.wait_stream(main)
            with torch.cuda.stream(s):
                e.root(obs=o, train=True)
                e.simulate(50)
        for s in streams:
            main.wait_stream(s)
        b.record()
        torch.cuda.synchronize()
        if it >= 3:
            ms.append(a.elapsed_time(b))
    ms.sort()
    return ms[len(ms) // 2]


for B in (512, 1024, 2048, 4096, 8192):
    e = make(B, 0)
    o = torch.randn(B, 4, device=dev)
    t = timed([e], [torch.cuda.Stream()], [o])
    print(f"single engine  B={B:6d}: {t:.3f} ms  -> {B * 50 / t / 1e3:.1f} M sims/s", flush=True)
    e.close()
for G in (2, 4):
    B = 4096 // G
    engs = [make(B, i * B) for i in range(G)]
    obs = [torch.randn(B, 4, device=dev) for _ in range(G)]
    t = timed(engs, [torch.cuda.Stream() for _ in range(G)], obs)
    print(f"{G} engines x {B} trees on {G} streams: {t:.3f} ms -> {4096 * 50 / t / 1e3:.1f} M sims/s", flush=True)
    for e in engs:
        e.close()
