"""TEST / BENCH INFRASTRUCTURE — the reference's CPU search path timed on the host cores.

Where a checkout of the reference is reachable ($SMZ_REFERENCE, baseline/_ref, /root/reference: the build
container) the UNMODIFIED reference is timed (kind "reference").  It is Python and does not exist on the GPU
box; there the timed thing is the oracle PORT (oracle/mcts_oracle.py + oracle/net_oracle.py: one tree, one
network call per node, batch 1, fp32 — the same structure as monte_carlo_tree_search.py:311-349 driving
muzero_model.py:802-909; kind "port", ~2.6x faster per core than the reference, i.e. a conservative baseline).  Workload = BASELINE.json configs[0]: CartPole MLP of
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


def _reference_worker(args):
    """The UNMODIFIED reference: monte_carlo_tree_search.py:311-349 driving a random-init `Muzero` MLP
    (muzero_model.py:802-909) — BASELINE.json configs[0], one single-threaded process per core."""
    seconds, seed = args
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = "1"
    import warnings
    import numpy as np
    import torch
    from oracle import ref_shim
    torch.set_num_threads(1)
    ref_mcts, _ = ref_shim.load()
    model = ref_shim.make_muzero(seed=seed)
    search = ref_mcts.Monte_carlo_tree_search(pb_c_base=19652, pb_c_init=1.25, discount=0.997,
                                              root_dirichlet_alpha=0.25, root_exploration_fraction=0.25,
                                              num_simulations=50, maxium_action_sample=2, number_of_player=1,
                                              custom_loop=None)
    np.random.seed(seed)
    g = np.random.default_rng(seed)

    def move():
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            search.run(observation=g.standard_normal((1, 4)).astype(np.float32), model=model, train=True)

    move()
    t0 = time.perf_counter()
    moves = 0
    while time.perf_counter() - t0 < seconds:
        move()
        moves += 1
    return moves, time.perf_counter() - t0


def _pool_rate(worker, n_procs, seconds):
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Pool(n_procs) as pool:
        res = pool.map(worker, [(seconds, 1000 + i) for i in range(n_procs)])
    return sum(50.0 * m / t for m, t in res), sum(m for m, _ in res)


def measure(n_procs=None, seconds=10.0, kind="auto"):
    """-> dict(value=sims/s aggregate, cores, kind, sample, per_core).  kind "reference" times the real reference
    when a checkout is reachable ($SMZ_REFERENCE, baseline/_ref, /root/reference — never on the GPU box), "port"
    the oracle port; "auto" prefers the reference and reports the port's rate next to it."""
    from oracle import ref_shim
    n_procs = n_procs or os.cpu_count() or 1
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ.setdefault(k, "1")
    what = "BASELINE configs[0] (CartPole MLP 450 shape, 1 tree/process, batch-1 fp32 network)"
    use_ref = kind == "reference" or (kind == "auto" and ref_shim.available())
    if use_ref and not ref_shim.available():
        raise RuntimeError("no reference checkout reachable")
    if use_ref:
        half = max(1.0, seconds / 2)
        rate, moves = _pool_rate(_reference_worker, n_procs, half)
        port_rate, _ = _pool_rate(_worker, n_procs, half)
        return {"value": rate, "unit": "sims/s", "cores": n_procs, "kind": "reference",
                "sample": f"{moves} moves x 50 simulations of {what} by the unmodified reference at "
                          f"{ref_shim.REFERENCE_DIR} in ~{half:.0f} s on {n_procs} single-threaded processes",
                "per_core": rate / n_procs, "port_value": port_rate}
    rate, moves = _pool_rate(_worker, n_procs, seconds)
    return {"value": rate, "unit": "sims/s", "cores": n_procs, "kind": "port",
            "sample": f"{moves} moves x 50 simulations of {what} in ~{seconds:.0f} s on {n_procs} "
                      f"single-threaded processes (oracle port; no reference checkout on this box)",
            "per_core": rate / n_procs}
