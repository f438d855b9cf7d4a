"""TEST / BENCH INFRASTRUCTURE — the reference's CPU search path timed on the host cores.

The reference is Python and does not exist on the GPU box, so the timed thing is the oracle PORT
(oracle/mcts_oracle.py + oracle/net_oracle.py: one tree, one network call per node, batch 1, fp32 —
the same structure and cost profile as monte_carlo_tree_search.py:311-349 driving
muzero_model.py:802-909).  Workload = BASELINE.json configs[0]: CartPole MLP of
config/experiment_450_config.json (obs 4, A 2, S 61, H 126, L 4), 50 simulations per move,
consecutive moves on synthetic N(0,1) observations, random-init weights (weights_init: N(0, 1/137)).
One single-threaded process per host core, each running moves for a bounded time.
"""
import os
import time


def _worker(args):
    seconds, seed = args
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = "1"
    import numpy as np
    from oracle import mcts_oracle as O
    from oracle import net_oracle as NO
    dims = (4, 2, 2, 61, 126, 4)
    _, total = NO.blob_layout(*dims)
    g = np.random.default_rng(seed)
    net = NO.NetOracle((g.standard_normal(total) / 137.035999).astype(np.float32), *dims)
    cfg = O.SearchConfig(pb_c_base=19652, pb_c_init=1.25, discount=0.997, root_dirichlet_alpha=0.25,
                         root_exploration_fraction=0.25, num_simulations=50, maxium_action_sample=2)
    rs = np.random.RandomState(seed)
    rng = O.MTUniforms(seed)

    def move():
        model = NO.NetModel(net, g.standard_normal(4))
        rng.log.clear()
        O.search(cfg, model, rng, train=True, dirichlet_fn=lambda n: rs.dirichlet([cfg.root_dirichlet_alpha] * n))

    move()                                   # warm-up
    t0 = time.perf_counter()
    moves = 0
    while time.perf_counter() - t0 < seconds:
        move()
        moves += 1
    return moves, time.perf_counter() - t0


def measure(n_procs=None, seconds=10.0):
    """-> dict(value=sims/s aggregate, cores, sample, per_core)."""
    import multiprocessing as mp
    n_procs = n_procs or os.cpu_count() or 1
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ.setdefault(k, "1")
    ctx = mp.get_context("spawn")
    with ctx.Pool(n_procs) as pool:
        res = pool.map(_worker, [(seconds, 1000 + i) for i in range(n_procs)])
    rate = sum(50.0 * m / t for m, t in res)
    return {"value": rate, "unit": "sims/s", "cores": n_procs, "kind": "port",
            "sample": f"{sum(m for m, _ in res)} moves x 50 simulations of BASELINE configs[0] (CartPole MLP 450 "
                      f"shape, 1 tree/process, batch-1 fp32 network) in ~{seconds:.0f} s on {n_procs} "
                      f"single-threaded processes",
            "per_core": rate / n_procs}
