"""bf16 tcgen05 network step bring-up: per-network error against the reference's fp32 outputs, then a
timed 4096 x 50 search in both network modes."""
import faulthandler
import os
import sys
import time

faulthandler.enable()
faulthandler.dump_traceback_later(150, exit=True)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import golden_io
from stochastic_muzero_b200 import ModelShape, SearchEngine

T0 = time.time()


def log(m):
    print(f"[{time.time() - T0:6.2f}s] {m}", flush=True)


SEARCH = dict(pb_c_base=19652, pb_c_init=1.25, discount=0.997, root_dirichlet_alpha=0.25,
              root_exploration_fraction=0.25, num_simulations=50, maxium_action_sample=2, number_of_player=1,
              custom_loop=None)
for name in ("mlp450_seed0", "ckpt450", "mlp_small", "mlp_l0"):
    z = golden_io.load_net_case(name)
    obs, A, C, S, H, L = [int(v) for v in z["dims"]]
    eng = SearchEngine(dict(SEARCH, maxium_action_sample=min(2, A)), A, C, max_trees=256,
                       model_shape=ModelShape(obs, A, C, S, H, L), net="bf16")
    eng.set_weights(z["weights"])
    torch.cuda.synchronize()
    log(f"{name}: bf16 engine ready")

    def err(a, b):
        return float(np.abs(a.cpu().numpy() - b).max())
    e = {}
    e["repr_h"] = err(eng.net_eval("repr", z["obs"])["hidden"], z["repr_h"])
    o = eng.net_eval("pred", z["repr_h"]); e["pred_policy"] = err(o["policy"], z["pred_policy"]); e["pred_value"] = err(o["value"], z["pred_value"])
    e["adyn_h"] = err(eng.net_eval("adyn", z["repr_h"], z["actions"])["hidden"], z["adyn_h"])
    o = eng.net_eval("apred", z["adyn_h"]); e["apred_policy"] = err(o["policy"], z["apred_policy"]); e["apred_value"] = err(o["value"], z["apred_value"])
    o = eng.net_eval("dyn", z["adyn_h"], z["actions"]); e["dyn_h"] = err(o["hidden"], z["dyn_h"]); e["dyn_reward"] = err(o["reward"], z["dyn_reward"])
    o = eng.net_eval("enc", z["obs"]); e["enc_probs"] = err(o["probs"], z["enc_probs"])
    log(f"{name}: max abs err vs reference fp32: " + ", ".join(f"{k}={v:.2e}" for k, v in e.items()))
    log(f"   scales: value~{np.abs(z['pred_value']).max():.2f} reward~{np.abs(z['dyn_reward']).max():.3f}")
    eng.close()

z = golden_io.load_net_case("mlp450_seed0")
shape = ModelShape(4, 2, 2, 61, 126, 4)
B, N = 4096, 50
obs = torch.randn(B, 4, generator=torch.Generator().manual_seed(0)).cuda()
res = {}
for net in ("fp32", "bf16"):
    eng = SearchEngine(SEARCH, 2, 2, max_trees=B, model_shape=shape, net=net, seed=7)
    eng.set_weights(z["weights"])
    for _ in range(3):
        eng.root(obs=obs, train=True); eng.simulate(N)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(10):
        eng.set_seed(100 + i)
        eng.root(obs=obs, train=True); eng.simulate(N)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    r = eng.read_roots()
    res[net] = r["visits"].cpu().numpy(), r["root_values"].cpu().numpy()
    log(f"{net}: {ms:.3f} ms per 4096x50 search = {B * N / ms / 1e3:.1f} M sims/s; stats {eng.stats()}")
    # per-kernel split
    eng.set_seed(5); eng.root(obs=obs, train=True); torch.cuda.synchronize()
    ts = [0.0, 0.0, 0.0]
    evs = []
    for s in range(N):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record(); eng.select(s); e[1].record(); eng.net_step(s); e[2].record(); eng.expand_backup(s); e[3].record()
        evs.append(e)
    torch.cuda.synchronize()
    for e in evs:
        for i in range(3):
            ts[i] += e[i].elapsed_time(e[i + 1]) / N
    log(f"{net}: per-sim kernels select {ts[0]*1e3:.1f} us, net {ts[1]*1e3:.1f} us, expand_backup {ts[2]*1e3:.1f} us")
    eng.close()
same = (res["fp32"][0] == res["bf16"][0]).all(1).mean()
log(f"identical root visit vectors fp32 vs bf16 (same seeds): {same:.3f}; mean |dvisit| "
    f"{np.abs(res['fp32'][0] - res['bf16'][0]).mean():.2f}; root value diff {np.abs(res['fp32'][1] - res['bf16'][1]).max():.3e}")
log("DONE")
