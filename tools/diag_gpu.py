"""Staged GPU bring-up with timestamps (writes progress immediately; used under gpurun)."""
import faulthandler
import os
import sys
import time

faulthandler.enable()
faulthandler.dump_traceback_later(100, exit=True)
T0 = time.time()


def log(msg):
    print(f"[{time.time() - T0:7.2f}s] {msg}", flush=True)


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
log("start")
import numpy as np
import torch
log(f"torch imported {torch.__version__}")
torch.cuda.init()
log(f"cuda init ok: {torch.cuda.get_device_name(0)}")
x = torch.zeros(4, device="cuda"); torch.cuda.synchronize()
log("first kernel ok")
from stochastic_muzero_b200 import ModelShape, SearchEngine
from stochastic_muzero_b200.weights import random_blob
log("package imported")
import golden_io
z = golden_io.load_tree_case("a2c2k2_n50_train")
c = z["config"]
B, N = len(z["n_nodes"]), c["num_simulations"]
eng = SearchEngine(c, 2, 2, max_trees=B, net="external", rng="tape")
log("engine (external, tape) created")
eng.set_uniform_tape(torch.from_numpy(z["uniforms"]))
eng.root(root_policy=torch.from_numpy(z["root_policy"]), root_to_play=torch.from_numpy(z["exp_root_to_play"]),
         train=True, dirichlet=torch.from_numpy(z["dirichlet"]))
torch.cuda.synchronize()
log("root ok")
pol, val, rew = (torch.from_numpy(z[k]).cuda() for k in ("sim_policy", "sim_value", "sim_reward"))
for s in range(N):
    slot, action, branch = eng.select(s)
    torch.cuda.synchronize()
    if s < 3:
        log(f"select {s} ok: {action.tolist()[:4]}")
    eng.expand_backup(s, pol[:, s].contiguous(), val[:, s].contiguous(), rew[:, s].contiguous())
    torch.cuda.synchronize()
    if s < 3:
        log(f"expand_backup {s} ok")
log("50 sims ok")
got = eng.export_tree(0)
golden_io.assert_dump_equal(got, golden_io.expected_dump(z, 0), "diag")
log("tree 0 bit-exact vs reference golden")
eng.close()

shape = ModelShape(4, 2, 2, 61, 126, 4)
search = dict(c)
eng = SearchEngine(search, 2, 2, max_trees=64, model_shape=shape, net="fp32", rng="philox", seed=1, record=True)
log("engine (fp32, philox) created")
eng.set_weights(random_blob(shape, 0))
torch.cuda.synchronize()
log("weights packed")
out = eng.net_eval("repr", np.zeros((8, 4), np.float32))
torch.cuda.synchronize()
log(f"net_eval ok {out['hidden'][0, :4].tolist()}")
eng.root(obs=torch.randn(64, 4), train=True)
torch.cuda.synchronize()
log("root (internal net + device dirichlet) ok")
slot, action, branch = eng.select(0); torch.cuda.synchronize(); log("select ok")
eng.net_step(0); torch.cuda.synchronize(); log("net_step ok")
eng.expand_backup(0); torch.cuda.synchronize(); log("expand_backup ok")
eng.root(obs=torch.randn(64, 4), train=True)
eng.simulate(50)
torch.cuda.synchronize()
log(f"simulate(50) via graph ok: {eng.stats()}")
t = time.time()
for _ in range(5):
    eng.root(obs=torch.randn(64, 4), train=True); eng.simulate(50)
torch.cuda.synchronize()
log(f"5 more searches: {(time.time() - t) / 5 * 1e3:.2f} ms each")
log("DONE")
