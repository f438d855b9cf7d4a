"""64-row-tile chain kernel (SMZ_M64=1) vs the 128-row plain kernel over a whole search: max differences."""
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
for m64 in (0, 1):
    os.environ.pop("SMZ_NO_PIPE", None); os.environ.pop("SMZ_M64", None)
    if m64: os.environ["SMZ_M64"] = "1"
    else: os.environ["SMZ_NO_PIPE"] = "1"
    eng = SearchEngine(SEARCH, 2, 2, max_trees=B, model_shape=ModelShape(4, 2, 2, 61, 126, 4), net="bf16", seed=99, record=True)
    eng.set_weights(z["weights"])
    eng.root(obs=obs, train=True)
    res = []
    for s in range(4):
        slot, act, br = eng.select(s)
        eng.net_step(s)
        torch.cuda.synchronize()
        h = eng.read_hidden(s + 1).cpu().numpy()
        eng.expand_backup(s)
        torch.cuda.synchronize()
        rec = {k: v.cpu().numpy() for k, v in eng.read_record().items()}
        res.append((br.cpu().numpy(), h, rec["sim_policy"][:, s].copy(), rec["sim_value"][:, s].copy(), rec["sim_reward"][:, s].copy()))
    out[m64] = res
    eng.close()
for s in range(4):
    a, b = out[0][s], out[1][s]
    print(f"sim {s}: branches equal {np.array_equal(a[0], b[0])}; hidden maxdiff {np.nanmax(np.abs(a[1]-b[1])):.3e} nan {np.isnan(b[1]).sum()}; "
          f"policy maxdiff {np.nanmax(np.abs(a[2]-b[2])):.3e} nan {np.isnan(b[2]).sum()}; value maxdiff {np.nanmax(np.abs(a[3]-b[3])):.3e}; "
          f"reward maxdiff {np.nanmax(np.abs(a[4]-b[4])):.3e}")
print("row0 hidden plain", out[0][0][1][0, :8]); print("row0 hidden m64  ", out[1][0][1][0, :8])
print("value plain", out[0][0][3][:4], "m64", out[1][0][3][:4]); print("policy plain", out[0][0][2][0], "m64", out[1][0][2][0])
