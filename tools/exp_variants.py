"""Whole-search time of cfg2-shaped searches under kernel-choice switches (environment variables read at engine
creation / launch).  CUDA events, median of 20 repetitions, L2 flushed between searches.
usage: exp_variants.py [net] [trees,trees,...] -- VAR=val,VAR=val  VAR=val ..."""
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
args = sys.argv[1:]
sep = args.index("--") if "--" in args else len(args)
net = args[0] if sep > 0 else "bf16"
sizes = [int(x) for x in args[1].split(",")] if sep > 1 else [4096]
variants = args[sep + 1:] or [""]
blob = random_blob(shape, seed=0)
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
SWITCHES = set()
for v in variants:
    for kv in filter(None, v.split(",")):
        SWITCHES.add(kv.split("=")[0])

for B in sizes:
    obs = torch.randn(B, 4, device=dev)
    for v in variants:
        for k in SWITCHES:
            os.environ.pop(k, None)
        for kv in filter(None, v.split(",")):
            k, val = kv.split("=")
            os.environ[k] = val
        e = SearchEngine(SEARCH, 2, 2, max_trees=B, model_shape=shape, net=net, rng="philox", seed=7)
        e.set_weights(blob)
        ms = []
        for it in range(24):
            flush.zero_()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            e.root(obs=obs, train=True)
            e.simulate(50)
            b.record()
            torch.cuda.synchronize()
            if it >= 4:
                ms.append(a.elapsed_time(b))
        ms.sort()
        t = ms[len(ms) // 2]
        vis = int(e.read_roots()["visits"].sum().item())
        print(f"B={B:6d} {v or 'default':40s}: {t:.3f} ms  {B * 50 / t / 1e3:7.1f} M sims/s  (visits {vis})", flush=True)
        e.close()
