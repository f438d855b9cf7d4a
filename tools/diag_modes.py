"""Per network-step mode (bf16 / f16 / tc32 / fp32): worst error against the reference's inference outputs
(tests/golden/net_*.npz), search-result agreement with the fp32 search on the same seeds, and search time."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, golden_io
from stochastic_muzero_b200 import ModelShape, SearchEngine
SEARCH = dict(pb_c_base=19652, pb_c_init=1.25, discount=0.997, root_dirichlet_alpha=0.25,
              root_exploration_fraction=0.25, num_simulations=50, maxium_action_sample=2, number_of_player=1, custom_loop=None)
MODES = sys.argv[1:] or ["fp32", "tc32", "f16", "bf16"]
for name in ("mlp450_seed0", "ckpt450"):
    z = golden_io.load_net_case(name)
    dims = [int(v) for v in z["dims"]]
    B = 4096
    obs = (torch.randn(B, 4, generator=torch.Generator().manual_seed(0)) * 0.1).cuda()
    res = {}
    for net in MODES:
        eng = SearchEngine(SEARCH, 2, 2, max_trees=B, model_shape=ModelShape(*dims), net=net, seed=11)
        eng.set_weights(z["weights"])
        err = {}
        err["repr_h"] = np.abs(eng.net_eval("repr", z["obs"])["hidden"].cpu().numpy() - z["repr_h"]).max()
        o = eng.net_eval("pred", z["repr_h"]); err["pred_pol"] = np.abs(o["policy"].cpu().numpy() - z["pred_policy"]).max()
        err["pred_val"] = np.abs(o["value"].cpu().numpy() - z["pred_value"]).max()
        err["adyn_h"] = np.abs(eng.net_eval("adyn", z["repr_h"], z["actions"])["hidden"].cpu().numpy() - z["adyn_h"]).max()
        o = eng.net_eval("apred", z["adyn_h"]); err["apred_val"] = np.abs(o["value"].cpu().numpy() - z["apred_value"]).max()
        o = eng.net_eval("dyn", z["adyn_h"], z["actions"]); err["dyn_h"] = np.abs(o["hidden"].cpu().numpy() - z["dyn_h"]).max()
        err["dyn_rew"] = np.abs(o["reward"].cpu().numpy() - z["dyn_reward"]).max()
        ts = []
        for it in range(6):
            eng.set_seed(11)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); eng.root(obs=obs, train=True); eng.simulate(50); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        r = eng.read_roots()
        res[net] = (r["visits"].cpu().numpy(), r["root_values"].cpu().numpy())
        print(f"{name} {net:5s}: {min(ts):.3f} ms/search ({B*50/min(ts)/1e3:.1f} M sims/s)  " +
              " ".join(f"{k}={v:.1e}" for k, v in err.items()), flush=True)
        eng.close()
    ref = res.get("fp32") or res[MODES[0]]
    for net in MODES:
        v32, v16 = ref[0], res[net][0]
        same = (v32 == v16).all(1).mean()
        argsame = (v32.argmax(1) == v16.argmax(1)).mean()
        rel = np.abs(ref[1] - res[net][1]) / np.maximum(np.abs(ref[1]), 1e-3)
        print(f"   {name} {net:5s} vs {'fp32' if 'fp32' in res else MODES[0]}: identical visit vectors {same:.4f}; same most-visited action {argsame:.4f}; "
              f"root value rel diff p99 {np.percentile(rel, 99):.2e}", flush=True)
