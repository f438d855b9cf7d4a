#!/usr/bin/env python
"""bench.py — MCTS simulations/sec of the batched Stochastic-MuZero search (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--net fp32|bf16]

A "step" is one whole batched search: root step (representation + prediction, root expansion,
Dirichlet mixing) + 50 simulations for 4096 concurrent trees (BASELINE.json configs[1]: CartPole MLP
of config/experiment_450_config.json, synthetic observations, random-init weights).  N > 1 shards
independent trees across ranks (4096 per GPU, weak scaling; the only collective is one NCCL weight
broadcast before timing).  One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEARCH = dict(pb_c_base=19652, pb_c_init=1.25, discount=0.997, root_dirichlet_alpha=0.25,
              root_exploration_fraction=0.25, num_simulations=50, maxium_action_sample=2,
              number_of_player=1, custom_loop=None)
DIMS = dict(obs_dim=4, action_dim=2, chance_dim=2, state_dim=61, hidden_dim=126, num_hidden_layers=4)
METRIC, UNIT = "mcts_simulations_per_sec", "sims/s"

# BASELINE.json configs.  cfg2 is the bench line (the metric is quoted on it); the others are parity-test
# shapes that can be timed on request (--workload) and are reported with the same JSON contract.
WORKLOADS = {
    "cfg2": dict(name="BASELINE configs[1]: CartPole MLP (obs 4, A 2, S 61, H 126, L 4)", trees=4096, sims=50,
                 dims=DIMS, K=2, net=None),
    "cfg3": dict(name="BASELINE configs[2]: stochastic 2048-like (4x4 board obs 16, 4 actions, 32 chance codes, K 32)",
                 trees=8192, sims=100, dims=dict(DIMS, obs_dim=16, action_dim=4, chance_dim=32), K=32, net=None),
    "cfg4": dict(name="BASELINE configs[3]: CartPole MLP, 65536 trees sharded across the ranks", trees=65536, sims=50,
                 dims=DIMS, K=2, net=None, total=True),
    "cfg5": dict(name="BASELINE configs[4]: vision ResNet-v2 (98x98 RGB, A 4, S 61, H 126, L 4)", trees=1024, sims=50,
                 dims=dict(action_dim=4, state_dim=61, hidden_dim=126, num_hidden_layers=4), K=2, net="vision"),
}


def vision_flops_per_sim(d):
    """2*MAC of the vision per-simulation networks: 3x3 convs on 3x7x7 maps + three 147->H->..->S/A MLP heads."""
    H, L, S, A = d["hidden_dim"], d["num_hidden_layers"], d["state_dim"], d["action_dim"]
    conv = lambda ci, co: 2 * ci * co * 9 * 49                                    # noqa: E731
    res = 3 * conv(3, 3)
    mlp = lambda n: 2 * (147 * H + L * H * H + H * n)                             # noqa: E731
    trunk = conv(4, 3) + L * res
    pred = L * res + 2 * 2 * 9 * 49 + mlp(S) + mlp(A)
    after = trunk + pred
    dyn = trunk + 2 * 12 * 49 + mlp(S) + pred
    return after, dyn, 0


def flops_per_sim(d):
    """2*MAC, trunk counted once, one-hot counted as dense S+A input."""
    S, H, L, A, C = d["state_dim"], d["hidden_dim"], d["num_hidden_layers"], d["action_dim"], d["chance_dim"]
    OH = max(A, C)
    after = 2 * ((S + OH) * H + L * H * H + H * S) + 2 * (S * H + L * H * H + H * C + H * S)
    dyn = 2 * ((S + OH) * H + L * H * H + 2 * H * S) + 2 * (S * H + L * H * H + H * A + H * S)
    root = 2 * (d["obs_dim"] * H + L * H * H + H * S) + 2 * (S * H + L * H * H + H * A + H * S)
    return after, dyn, root


def tree_bytes_per_sim(depth, K, S):
    """SURVEY.md §8d algorithmic bytes of the tree kernels per simulation (select + expand + backup)."""
    return depth * (16 * K + 8) + 8 * S + 4 + 12 * K + 20 * (depth + 1) + 8


class ClockSampler:
    FIELDS = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(device)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self, t0=None, t1=None):
        """Median SM clock and throttle reasons over the samples taken inside [t0, t1] (epoch seconds)."""
        import datetime
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        def collect(windowed):
            sm, mx, reasons = [], [], set()
            for line in out.strip().splitlines():
                f = [x.strip() for x in line.split(",")]
                if len(f) < 8:
                    continue
                try:
                    ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    if windowed and t0 is not None and not (t0 - 0.05 <= ts <= t1 + 0.05):
                        continue
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            return sorted(sm), mx, reasons
        sm, mx, reasons = collect(True)
        window = "timed region"
        if not sm:      # a very short timed region can fall between two 50 ms samples: use the whole run (warm-up included)
            sm, mx, reasons = collect(False)
            window = "whole run (timed region shorter than the sampling period)"
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def run_reference(args):
    """Reference arm: the reference's CPU search path (oracle port, see oracle/cpu_baseline.py) on all
    host cores; each step is a bounded sample of configs[0]."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_baseline
    per_step = max(2.0, min(10.0, 60.0 / max(1, args.steps + args.warmup)))
    for _ in range(min(args.warmup, 1)):
        cpu_baseline.measure(seconds=1.0)
    rates, t0 = [], time.perf_counter()
    for _ in range(args.steps):
        rates.append(cpu_baseline.measure(seconds=per_step))
    wall = time.perf_counter() - t0
    value = sum(r["value"] for r in rates) / len(rates)
    base = dict(rates[-1], value=value)
    base.pop("per_core", None)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE configs[0]: CartPole MLP 450 shape, 1 tree per host process, 50 "
                                   "simulations/move, reference CPU algorithm (oracle port; the Python reference "
                                   "cannot travel to the GPU box)", "sims_per_move": 50},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--net", default=os.environ.get("SMZ_BENCH_NET", "auto"), choices=["auto", "fp32", "bf16"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--trees", type=int, default=None, help="concurrent trees per GPU (default: the workload's)")
    ap.add_argument("--sims", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--profile-only", action="store_true", help="timed steps only (for runs under ncu)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import torch.distributed as dist
    from stochastic_muzero_b200 import (ModelShape, Monte_carlo_tree_search, PackedModel, SearchEngine, VisionShape,
                                        random_blob, vision_blob_layout)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    wl = WORKLOADS[args.workload]
    dims = wl["dims"]
    N = args.sims or wl["sims"]
    B = args.trees or (wl["trees"] // world if wl.get("total") else wl["trees"])
    search = dict(SEARCH, num_simulations=N, maxium_action_sample=wl["K"])
    vision = wl["net"] == "vision"
    net = "vision" if vision else args.net
    if net == "auto":
        net = os.environ.get("SMZ_DEFAULT_NET", "bf16")
    shape = VisionShape(**dims) if vision else ModelShape(**dims)
    A = shape.action_dim
    C = A if vision else shape.chance_dim
    obs_shape = (3 * 98 * 98,) if vision else (shape.obs_dim,)
    eng = SearchEngine(search, A, C, max_trees=B, model_shape=shape, net=net,
                       rng="philox", seed=20240 + rank, tree_id_offset=rank * B, device=local)

    # weights: rank 0 draws them, one NCCL broadcast hands them to the other shards
    if vision:      # torch default init is not reproducible here: N(0, 0.05) convs / heads, identity-like BatchNorm
        lay, tot = vision_blob_layout(shape)
        w0 = (np.random.default_rng(0).standard_normal(tot) * 0.05).astype(np.float32)
        for k, (o, shp) in lay.items():
            if k.endswith(".bn"):
                c = shp[1]
                w0[o:o + 4 * c] = np.concatenate([np.ones(c), np.zeros(c), np.zeros(c), np.ones(c)])
    else:
        w0 = random_blob(shape, seed=0)
    blob = torch.from_numpy(w0).to(dev) if rank == 0 else \
        torch.empty(eng.dims.weight_blob_floats, dtype=torch.float32, device=dev)
    bcast_ms = 0.0
    sampler = ClockSampler(local) if rank == 0 else None     # started early: nvidia-smi takes a while to spin up
    if world > 1:
        dist.all_reduce(torch.zeros(1, device=dev))           # NCCL communicator set-up is not the broadcast
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dist.broadcast(blob, src=0)
        torch.cuda.synchronize()
        bcast_ms = 1e3 * (time.perf_counter() - t0)
    eng.set_weights(blob)
    gen = torch.Generator().manual_seed(rank)
    if args.workload == "cfg3":      # 4x4 board, log2(tile)/16 with tiles drawn from {0..11}
        obs = (torch.randint(0, 12, (B,) + obs_shape, generator=gen).float() / 16.0).to(dev)
    elif vision:
        obs = torch.rand((B,) + obs_shape, generator=gen).to(dev)
    else:
        obs = torch.randn((B,) + obs_shape, generator=gen).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step():
        eng.root(obs=obs, train=True)
        eng.simulate(N)

    for i in range(args.warmup):
        eng.set_seed(1000 + i, rank * B)
        step()
    torch.cuda.synchronize()
    launches0 = None
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    wall0 = time.time()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches = 0
    for i, (a, b) in enumerate(evs):
        eng.set_seed(2000 + i, rank * B)
        flush.zero_()                      # evict the arena from L2 between timed steps (not timed)
        a.record()
        step()
        b.record()
        launches += 2 * N + 1 + 3     # select + N x (net, tree) + root net, dirichlet, root expand
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop(wall0, time.time()) if sampler else None
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    stats = eng.stats()
    value = world * B * N * args.steps / (dev_ms * 1e-3)

    if args.profile_only:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "ms_per_step": dev_ms / args.steps,
                              "note": "profile-only run; not a bench value if taken under ncu"}), flush=True)
        return

    # ---- roofline pass: per-kernel CUDA-event timing of one more search, step by step ----------------
    eng.set_seed(3000, rank * B)
    eng.root(obs=obs, train=True)
    torch.cuda.synchronize()
    ev = lambda: torch.cuda.Event(enable_timing=True)   # noqa: E731
    marks, n_dyn = [], torch.zeros((), dtype=torch.int64, device=dev)
    NET_REP = 8
    for s in range(N):
        e0, e1, e2, e3 = ev(), ev(), ev(), ev()
        e0.record(); _, _, br = eng.select(s); e1.record()
        for _ in range(NET_REP):            # the network step is idempotent: back-to-back launches hide the host launch gap
            eng.net_step(s)
        e2.record(); eng.expand_backup(s); e3.record()
        n_dyn += br.sum()
        marks.append((e0, e1, e2, e3))
    torch.cuda.synchronize()
    t_sel = sum(m[0].elapsed_time(m[1]) for m in marks) / N
    t_net = sum(m[1].elapsed_time(m[2]) for m in marks) / (N * NET_REP)
    t_exp = sum(m[2].elapsed_time(m[3]) for m in marks) / N
    f_after, f_dyn, f_root = vision_flops_per_sim(dims) if vision else flops_per_sim(dims)
    n_dyn = int(n_dyn.item())
    flops_launch = ((B * N - n_dyn) * f_after + n_dyn * f_dyn) / N
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tensor_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    which = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    achieved_tf = flops_launch / (t_net * 1e-3) / 1e12
    depth = stats["mean_leaf_depth"]
    tb = tree_bytes_per_sim(depth, min(wl["K"], max(A, C)), 147 if vision else dims["state_dim"]) * B
    achieved_gbs = tb / ((t_sel + t_exp) * 1e-3) / 1e9
    kname = {"bf16": "k_bf16_chain_m64 / k_bf16_chain_pipe, tcgen05 bf16", "fp32": "k_net_sim, fp32 CUDA cores",
             "vision": "k_vision_step, fp32 CUDA cores"}[net]
    roofline = {"kernel": "network step (%s)" % kname, "bound": "tensor", "achieved": achieved_tf,
                "peak": tensor_peak, "unit": "TFLOP/s", "frac": achieved_tf / tensor_peak, "traffic": None,
                "peak_source": which + ", sustained bf16", "avg_launch_us": 1e3 * t_net,
                "algorithmic_flops_per_launch": flops_launch,
                "how": "CUDA events around %d back-to-back launches of the (idempotent) network step of every simulation of "
                       "one extra search run step by step on the launching stream after the timed region" % NET_REP,
                # share of the real (graph + PDL) step: launches x per-launch time / measured step time.  Kernels overlap
                # a little under PDL, so the shares of all kernels add up to slightly more than 1.
                "share_of_step": min(1.0, N * t_net / (dev_ms / args.steps))}
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if args.workload == "cfg2" and net in tr.get("network_step", {}):
            roofline["traffic"] = tr["network_step"][net]
            roofline["traffic_source"] = tr.get("source")
    except Exception:
        tr = {}
    roofline_tree = {"kernel": "k_select + k_expand_backup", "bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak,
                     "unit": "GB/s", "frac": achieved_gbs / hbm_peak, "traffic": None, "peak_source": which,
                     "avg_launch_us": {"select": 1e3 * t_sel, "expand_backup": 1e3 * t_exp},
                     "algorithmic_bytes_per_launch_pair": tb, "mean_leaf_depth": depth}

    # ---- end to end through the public API: host observations in, host visit counts / values out -----
    mcts = Monte_carlo_tree_search(**{k: search[k] for k in search}, net=net, device=local, seed=77 + rank, max_batch=B)
    model = PackedModel(blob.cpu().numpy(), shape)
    obs_host = obs.cpu().pin_memory()
    for _ in range(3):
        r = mcts.run_batch(obs_host, model, train=True)
        r.visit_counts.cpu(); r.root_values.cpu()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = mcts.run_batch(obs_host, model, train=True)
        v_host, rv_host = r.visit_counts.cpu(), r.root_values.cpu()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    assert int(v_host.sum()) == B * N
    e2e = {"value": world * B * N * args.steps / e2e_s, "unit": UNIT,
           "h2d_bytes_per_step": int(obs_host.numel() * 4),
           "d2h_bytes_per_step": int(v_host.numel() * 4 + rv_host.numel() * 4),
           "ms_per_step": 1e3 * e2e_s / args.steps,
           "api": "Monte_carlo_tree_search.run_batch(pinned host observations) -> visit counts + root values on host"}

    # ---- the reference's own call, one tree at a time (BASELINE configs[0] shape through the drop-in run()) ----
    single = None
    if rank == 0 and world == 1 and not vision:
        one = Monte_carlo_tree_search(**{k: search[k] for k in search}, net="fp32", device=local, seed=5)
        for _ in range(3):
            one.run(observation=torch.randn(1, shape.obs_dim), model=model, train=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        moves = 20
        for _ in range(moves):
            root = one.run(observation=torch.randn(1, shape.obs_dim), model=model, train=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        single = {"ms_per_move": 1e3 * dt / moves, "value": moves * N / dt, "unit": UNIT,
                  "api": "Monte_carlo_tree_search.run(observation, model, train) -> Node, fp32 network step, 1 tree"}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import cpu_baseline
        cpu = cpu_baseline.measure(seconds=args.cpu_seconds)
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16" if net == "bf16" else "f32", "data": "synthetic",
                "config": {"workload": f"{wl['name']}: {B} concurrent trees x {N} simulations per GPU, synthetic "
                                       f"observations, random-init weights, device Philox RNG",
                           "workload_id": args.workload,
                           "trees_per_gpu": B, "simulations": N, "network_step": net, "tree_arithmetic": "f32/f64 "
                           "(reference numpy semantics)", "l2": "256 MiB flush buffer written between timed steps",
                           "sharding": f"{world} x {B} independent trees, no data-path collective",
                           "weight_broadcast_ms": bcast_ms},
                "roofline": roofline, "roofline_tree": roofline_tree, "cpu_baseline": cpu, "e2e": e2e,
                "gpu_launches": launches, "clocks": clocks, "mean_leaf_depth": depth, "single_tree_dropin": single}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
