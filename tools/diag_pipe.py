import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, golden_io
from stochastic_muzero_b200 import ModelShape, SearchEngine
SEARCH = dict(pb_c_base=19652, pb_c_init=1.25, discount=0.997, root_dirichlet_alpha=0.25,
              root_exploration_fraction=0.25, num_simulations=50, maxium_action_sample=2, number_of_player=1, custom_loop=None)
z = golden_io.load_net_case("ckpt450")
B = 300
obs = torch.randn(B, 4, generator=torch.Generator().manual_seed(1)).cuda()
out = {}
for pipe in (0, 1):
    if pipe: os.environ.pop("SMZ_NO_PIPE", None)
    else: os.environ["SMZ_NO_PIPE"] = "1"
    eng = SearchEngine(SEARCH, 2, 2, max_trees=B, model_shape=ModelShape(4, 2, 2, 61, 126, 4), net="bf16", seed=3, record=True)
    eng.set_weights(z["weights"])
    eng.root(obs=obs, train=True)
    res = []
    for s in range(3):
        slot, act, br = eng.select(s)
        eng.net_step(s)
        torch.cuda.synchronize()
        h = eng.read_hidden(s + 1).cpu().numpy()
        eng.expand_backup(s)
        torch.cuda.synchronize()
        rec = {k: v.cpu().numpy() for k, v in eng.read_record().items()}
        res.append((br.cpu().numpy(), h, rec["sim_policy"][:, s].copy(), rec["sim_value"][:, s].copy(), rec["sim_reward"][:, s].copy()))
    out[pipe] = res
    eng.close()
for s in range(3):
    a, b = out[0][s], out[1][s]
    print(f"sim {s}: branches equal {np.array_equal(a[0], b[0])}; hidden maxdiff {np.nanmax(np.abs(a[1]-b[1])):.3e} nan {np.isnan(b[1]).sum()}; "
          f"policy maxdiff {np.nanmax(np.abs(a[2]-b[2])):.3e} nan {np.isnan(b[2]).sum()}; value maxdiff {np.nanmax(np.abs(a[3]-b[3])):.3e} nan {np.isnan(b[3]).sum()}; "
          f"reward maxdiff {np.nanmax(np.abs(a[4]-b[4])):.3e}")
    bad = np.flatnonzero(np.abs(a[1]-b[1]).max(1) > 1e-3)
    print("   rows with hidden diff:", bad[:20], "count", len(bad), " cols:", np.flatnonzero(np.abs(a[1]-b[1]).max(0) > 1e-3)[:40])
for s in range(2):
    a, b = out[0][s], out[1][s]
    print("sim", s, "hidden shape", b[1].shape, "nan cols(pipe):", np.unique(np.argwhere(np.isnan(b[1]))[:, 1])[:20], "nan(plain):", np.isnan(a[1]).sum())
    d = np.abs(a[1] - b[1]); d[np.isnan(d)] = 0
    print("  worst cols by maxdiff:", np.argsort(-d.max(0))[:10], np.sort(-d.max(0))[:10] * -1)
    print("  row0 plain:", a[1][0, :12]); print("  row0 pipe :", b[1][0, :12])
    print("  policy row0", a[2][0], b[2][0], "value", a[3][:4], b[3][:4])
print("nan rows sim0:", np.flatnonzero(np.isnan(out[1][0][1]).any(1)), "value0 rows:", np.flatnonzero(out[1][0][3]==0)[:20])
