"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by RUNNING THE REFERENCE in this container.

    python oracle/make_golden.py            # needs /root/reference (or $SMZ_REFERENCE)

For every case the reference's own ``Monte_carlo_tree_search.run`` is executed with
  * ``np.random.{choice,uniform,dirichlet}`` routed through a private ``RandomState`` whose every
    underlying ``random_sample()`` double is logged in consumption order (the draw count of a call is
    found by advancing a clone of the pre-call state until it equals the post-call state), and
  * a recording wrapper around the model (a scripted stub, or the reference ``Muzero`` MLP) logging
    each ``*_inference`` result,
and the resulting tree (every node's visit_count / value_sum / reward / prior / is_chance / to_play,
the key path of every simulation, min-max bounds) is dumped next to the tape.  The committed files
are what pins ``oracle/mcts_oracle.py`` / ``oracle/net_oracle.py`` and, on the GPU box (where the
reference does not exist), what the CUDA engine is compared with.
"""
import contextlib
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shim  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


# --------------------------------------------------------------------------------------------------
# RNG recorder
# --------------------------------------------------------------------------------------------------
class RecordingRandom:
    def __init__(self, seed):
        self.rs = np.random.RandomState(seed)
        self.uniforms = []
        self.dirichlet = None

    def _clone(self):
        c = np.random.RandomState()
        c.set_state(self.rs.get_state())
        return c

    @staticmethod
    def _same(a, b):
        sa, sb = a.get_state(), b.get_state()
        return sa[2] == sb[2] and np.array_equal(sa[1], sb[1])

    def _log_until(self, clone):
        for _ in range(100000):
            if self._same(clone, self.rs):
                return
            self.uniforms.append(float(clone.random_sample()))
        raise RuntimeError("could not align RNG state")

    def choice(self, a, size=None, replace=True, p=None):
        clone = self._clone()
        out = self.rs.choice(a, size=size, replace=replace, p=p)
        self._log_until(clone)
        return out

    def uniform(self, low=0.0, high=1.0, size=None):
        clone = self._clone()
        out = self.rs.uniform(low=low, high=high, size=size)
        n0 = len(self.uniforms)
        self._log_until(clone)
        u = np.array(self.uniforms[n0:])
        assert np.array_equal(np.atleast_1d(out), low + (high - low) * u), "uniform != low+(high-low)*u"
        return out

    def dirichlet_(self, alpha, size=None):
        out = self.rs.dirichlet(alpha, size)
        assert self.dirichlet is None
        self.dirichlet = np.array(out, dtype=np.float64)
        return out

    @contextlib.contextmanager
    def patched(self):
        saved = (np.random.choice, np.random.uniform, np.random.dirichlet)
        np.random.choice, np.random.uniform, np.random.dirichlet = self.choice, self.uniform, self.dirichlet_
        try:
            yield self
        finally:
            np.random.choice, np.random.uniform, np.random.dirichlet = saved


# --------------------------------------------------------------------------------------------------
# models
# --------------------------------------------------------------------------------------------------
class StubModel:
    """Scripted model: random softmax policies of width A (prediction) / C (afterstate prediction),
    values and rewards from a seeded generator.  Exposes the five inference methods the reference
    search calls (monte_carlo_tree_search.py:182, :198, :271, :276, :280, :285)."""

    def __init__(self, action_dim, chance_dim, seed, value_scale=1.0, peaky=1.0):
        self.A, self.C = action_dim, chance_dim
        self.g = np.random.default_rng(seed)
        self.value_scale, self.peaky = value_scale, peaky

    def _policy(self, n):
        z = self.g.normal(size=n) * self.peaky
        e = np.exp(z - z.max())
        return (e / e.sum()).astype(np.float32)[None, :]

    def _scalar(self):
        return np.float32(self.g.normal() * self.value_scale)

    def representation_function_inference(self, obs):
        return torch.zeros(1, 1)

    def prediction_function_inference(self, h):
        return self._policy(self.A), self._scalar()

    def afterstate_prediction_function_inference(self, h):
        return self._policy(self.C), self._scalar()

    def afterstate_dynamics_function_inference(self, h, a):
        return torch.zeros(1, 1)

    def dynamics_function_inference(self, h, a):
        return self._scalar(), torch.zeros(1, 1)


class RecordingModel:
    def __init__(self, inner):
        self.inner = inner
        self.calls = []

    def representation_function_inference(self, obs):
        h = self.inner.representation_function_inference(obs)
        self.calls.append(("repr", h))
        return h

    def prediction_function_inference(self, h):
        policy, value = self.inner.prediction_function_inference(h)
        self.calls.append(("pred", np.array(policy[0], dtype=np.float32), np.float32(value)))
        return policy, value

    def afterstate_prediction_function_inference(self, h):
        policy, value = self.inner.afterstate_prediction_function_inference(h)
        self.calls.append(("apred", np.array(policy[0], dtype=np.float32), np.float32(value)))
        return policy, value

    def afterstate_dynamics_function_inference(self, h, a):
        out = self.inner.afterstate_dynamics_function_inference(h, a)
        self.calls.append(("adyn", int(a), out))
        return out

    def dynamics_function_inference(self, h, a):
        reward, out = self.inner.dynamics_function_inference(h, a)
        self.calls.append(("dyn", int(a), np.float32(reward), out))
        return reward, out


# --------------------------------------------------------------------------------------------------
# one recorded reference run
# --------------------------------------------------------------------------------------------------
def dump_reference_tree(root, min_max_stats):
    cols = {k: [] for k in ("depth", "key", "visit", "value_sum", "reward", "prior", "is_chance", "to_play",
                            "expanded")}
    types = set()

    def walk(node, key, depth):
        cols["depth"].append(depth)
        cols["key"].append(int(key))
        cols["visit"].append(int(node.visit_count))
        cols["value_sum"].append(np.float32(node.value_sum))
        cols["reward"].append(np.float32(node.reward))
        cols["prior"].append(np.float64(node.prior))
        cols["is_chance"].append(bool(node.is_chance))
        cols["to_play"].append(int(node.to_play))
        cols["expanded"].append(node.expanded())
        types.add((type(node.value_sum).__name__, type(node.prior).__name__, type(node.reward).__name__))
        keys = list(node.children.keys())
        assert keys == sorted(keys)
        for k in keys:
            walk(node.children[k], k, depth + 1)

    walk(root, -1, 0)
    out = {
        "depth": np.array(cols["depth"], np.int32), "key": np.array(cols["key"], np.int32),
        "visit": np.array(cols["visit"], np.int32), "value_sum": np.array(cols["value_sum"], np.float32),
        "reward": np.array(cols["reward"], np.float32), "prior": np.array(cols["prior"], np.float64),
        "is_chance": np.array(cols["is_chance"], np.int8), "to_play": np.array(cols["to_play"], np.int32),
        "expanded": np.array(cols["expanded"], np.int8),
        "minmax": np.array([min_max_stats.minimum, min_max_stats.maximum], np.float32),
    }
    return out, types


READOUT_TEMPERATURES = (0.0, 0.0625, 0.2, 0.25, 0.3, 0.5, 0.7, 1.0, 2.0, 3.0)


def reference_readout(root, seed):
    """The reference's own read-out of a searched root (game.py:179-216) for a set of temperatures, with the
    uniform behind np.random.choice recorded.  Uses Game's methods on a bare instance."""
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sys.path.insert(0, ref_shim.REFERENCE_DIR)
        try:
            import game as ref_game
        finally:
            sys.path.remove(ref_shim.REFERENCE_DIR)
    g = ref_game.Game.__new__(ref_game.Game)
    g.child_visits, g.root_values = [], []
    g.store_search_statistics(root)
    out = {"stored": np.asarray(g.child_visits[0], np.float64), "root_value": np.float32(g.root_values[0]),
           "policy": [], "index": [], "u": [], "sampled": []}
    keys = list(root.children.keys())
    for t_i, T in enumerate(READOUT_TEMPERATURES):
        action, policy, _reward = g.policy_action_reward_from_tree(root)
        policy = g.softmax_stable(policy, temperature=T)
        rr = RecordingRandom(seed * 131 + t_i)
        with rr.patched():
            sel = g.select_action(action, policy, T)
        out["policy"].append(np.asarray(policy, np.float64))
        out["index"].append(keys.index(sel))
        out["u"].append(rr.uniforms[0] if rr.uniforms else 0.5)
        out["sampled"].append(len(rr.uniforms) > 0)
    return out


def record_run(mcts_kwargs, model, seed, train=True, obs=None, prior_runs=0):
    """Run the reference search once and return (tape, expected) dicts."""
    ref_mcts, _ = ref_shim.load()
    mcts = ref_mcts.Monte_carlo_tree_search(**mcts_kwargs)
    for _ in range(prior_runs):          # advance Player_cycle.global_count like earlier moves would
        mcts.cycle.global_step()
    # trace the keys chosen per simulation by wrapping select_child
    paths, current = [], []
    orig_select = mcts.select_child
    orig_init = mcts.initialize_history_node_searchpath_variable

    def select_child():
        a, c = orig_select()
        current.append(int(a))
        return a, c

    def init_vars():
        if current:
            paths.append(list(current))
            current.clear()
        return orig_init()

    mcts.select_child, mcts.initialize_history_node_searchpath_variable = select_child, init_vars
    rec_model = RecordingModel(model)
    rr = RecordingRandom(seed)
    with rr.patched(), torch.no_grad():
        root = mcts.run(observation=obs, model=rec_model, train=train)
    if current:
        paths.append(list(current))
    expected, types = dump_reference_tree(root, mcts.min_max_stats)
    expected["root_to_play"] = np.int32(root.to_play)

    # tape: root policy, then per simulation (branch, policy, value, reward)
    calls = rec_model.calls
    assert calls[0][0] == "repr" and calls[1][0] == "pred"
    root_policy, root_value = calls[1][1], calls[1][2]
    N = mcts_kwargs["num_simulations"]
    sims = calls[2:]
    assert len(sims) == 2 * N
    W = max([root_policy.shape[0]] + [c[1].shape[0] for c in sims if c[0] in ("pred", "apred")])
    sim_policy = np.zeros((N, W), np.float32)
    sim_width = np.zeros(N, np.int32)
    sim_value = np.zeros(N, np.float32)
    sim_reward = np.zeros(N, np.float32)
    sim_branch = np.zeros(N, np.int8)
    sim_action = np.zeros(N, np.int32)
    hidden = []
    for s in range(N):
        first, second = sims[2 * s], sims[2 * s + 1]
        if first[0] == "dyn":
            assert second[0] == "pred"
            sim_branch[s], sim_action[s], sim_reward[s] = 1, first[1], first[2]
            hidden.append(first[3])
        else:
            assert first[0] == "adyn" and second[0] == "apred"
            sim_branch[s], sim_action[s] = 0, first[1]
            hidden.append(first[2])
        w = second[1].shape[0]
        sim_policy[s, :w], sim_width[s], sim_value[s] = second[1], w, second[2]
    tape = {
        "uniforms": np.array(rr.uniforms, np.float64),
        "dirichlet": rr.dirichlet if rr.dirichlet is not None else np.zeros(0, np.float64),
        "root_policy": root_policy, "root_value": np.float32(root_value),
        "sim_policy": sim_policy, "sim_width": sim_width, "sim_value": sim_value, "sim_reward": sim_reward,
        "sim_branch": sim_branch, "sim_action": sim_action,
    }
    expected["paths"] = paths
    expected["readout"] = reference_readout(root, seed)
    extra = {"root_hidden": calls[0][1], "sim_hidden": hidden, "types": sorted(types)}
    return tape, expected, extra


def pack_batch(cfg, train, runs):
    """Stack B recorded runs (ragged) into rectangular arrays + counts."""
    B = len(runs)
    N = cfg["num_simulations"]
    out = {"config_json": np.array(json.dumps(cfg)), "train": np.int8(train)}
    Umax = max(len(t["uniforms"]) for t, _ in runs)
    W = max(max(t["sim_policy"].shape[1] if N else 0, t["root_policy"].shape[0]) for t, _ in runs)
    A = runs[0][0]["root_policy"].shape[0]
    Mmax = max(len(e["visit"]) for _, e in runs)
    Lmax = max([1] + [len(p) for _, e in runs for p in e["paths"]])
    out["uniforms"] = np.zeros((B, Umax), np.float64)
    out["n_uniforms"] = np.zeros(B, np.int32)
    out["dirichlet"] = np.zeros((B, A), np.float64)
    out["root_policy"] = np.zeros((B, A), np.float32)
    out["sim_policy"] = np.zeros((B, N, W), np.float32)
    for k, dt in (("sim_width", np.int32), ("sim_value", np.float32), ("sim_reward", np.float32),
                  ("sim_branch", np.int8), ("sim_action", np.int32)):
        out[k] = np.zeros((B, N), dt)
    out["n_nodes"] = np.zeros(B, np.int32)
    for k, dt in (("depth", np.int32), ("key", np.int32), ("visit", np.int32), ("value_sum", np.float32),
                  ("reward", np.float32), ("prior", np.float64), ("is_chance", np.int8),
                  ("to_play", np.int32), ("expanded", np.int8)):
        out["exp_" + k] = np.zeros((B, Mmax), dt)
    out["exp_minmax"] = np.zeros((B, 2), np.float32)
    out["exp_root_to_play"] = np.zeros(B, np.int32)
    out["exp_paths"] = np.full((B, N, Lmax), -1, np.int32)
    nT = len(READOUT_TEMPERATURES)
    out["readout_temperature"] = np.array(READOUT_TEMPERATURES, np.float64)
    out["readout_stored"] = np.zeros((B, A), np.float64)
    out["readout_root_value"] = np.zeros(B, np.float32)
    out["readout_policy"] = np.zeros((B, nT, A), np.float64)
    out["readout_index"] = np.zeros((B, nT), np.int32)
    out["readout_u"] = np.zeros((B, nT), np.float64)
    out["readout_sampled"] = np.zeros((B, nT), np.int8)
    for b, (t, e) in enumerate(runs):
        u = t["uniforms"]
        out["uniforms"][b, :len(u)] = u
        out["n_uniforms"][b] = len(u)
        if len(t["dirichlet"]):
            out["dirichlet"][b] = t["dirichlet"]
        out["root_policy"][b] = t["root_policy"]
        if N:
            w = t["sim_policy"].shape[1]
            out["sim_policy"][b, :, :w] = t["sim_policy"]
        for k in ("sim_width", "sim_value", "sim_reward", "sim_branch", "sim_action"):
            out[k][b] = t[k]
        m = len(e["visit"])
        out["n_nodes"][b] = m
        for k in ("depth", "key", "visit", "value_sum", "reward", "prior", "is_chance", "to_play", "expanded"):
            out["exp_" + k][b, :m] = e[k]
        out["exp_minmax"][b] = e["minmax"]
        out["exp_root_to_play"][b] = e["root_to_play"]
        for s, p in enumerate(e["paths"]):
            out["exp_paths"][b, s, :len(p)] = p
        ro = e["readout"]
        out["readout_stored"][b], out["readout_root_value"][b] = ro["stored"], ro["root_value"]
        out["readout_policy"][b] = np.stack(ro["policy"])
        out["readout_index"][b], out["readout_u"][b], out["readout_sampled"][b] = ro["index"], ro["u"], ro["sampled"]
    return out


def stub_case(name, A, C, K, N, train, B, players=1, custom_loop=None, discount=0.997, seed0=0,
              value_scale=1.0, peaky=1.0, prior_runs=0, alpha=0.25, frac=0.25, pb_c_base=19652, pb_c_init=1.25):
    cfg = dict(pb_c_base=pb_c_base, pb_c_init=pb_c_init, discount=discount, root_dirichlet_alpha=alpha,
               root_exploration_fraction=frac, num_simulations=N, maxium_action_sample=K,
               number_of_player=players, custom_loop=custom_loop)
    runs = []
    for b in range(B):
        model = StubModel(A, C, seed=seed0 + 1000 + b, value_scale=value_scale, peaky=peaky)
        tape, exp, _ = record_run(cfg, model, seed=seed0 + b, train=train, obs=torch.zeros(1, 1),
                                  prior_runs=(prior_runs + b) % max(1, players if custom_loop is None
                                                                   else len(custom_loop.split(">"))))
        runs.append((tape, exp))
    cfg_out = dict(cfg, action_dim=A, chance_dim=C)
    np.savez_compressed(os.path.join(GOLDEN_DIR, f"tree_{name}.npz"), **pack_batch(cfg_out, train, runs))
    print(f"tree_{name}: B={B} nodes<= {max(len(e['visit']) for _, e in runs)} "
          f"uniforms<= {max(len(t['uniforms']) for t, _ in runs)}")


# --------------------------------------------------------------------------------------------------
# packed weights (the engine's hand-off format, include/smz.h "weight blob") from a reference Muzero
# --------------------------------------------------------------------------------------------------
def pack_reference_weights(mz):
    """fp32 blob in the order documented in include/smz.h; tied linear_mid packed once
    (neural_network_mlp_model.py:31-37)."""
    L = mz.number_of_hidden_layer

    def seq(module_seq, out_index=-1):
        mods = [m for m in module_seq if isinstance(m, torch.nn.Linear)]
        parts = [mods[0].weight, mods[0].bias]
        if L > 0:
            assert all(m is mods[1] for m in mods[1:1 + L])
            parts += [mods[1].weight, mods[1].bias]
        return parts, mods[-1]

    blob = []
    p, out = seq(mz.representation_function.state_norm); blob += p + [out.weight, out.bias]
    p, pol = seq(mz.prediction_function.policy); _, val = seq(mz.prediction_function.value)
    blob += p + [pol.weight, pol.bias, val.weight, val.bias]
    p, st = seq(mz.afterstate_dynamics_function.next_state_normalized); blob += p + [st.weight, st.bias]
    p, pol = seq(mz.afterstate_prediction_function.policy); _, val = seq(mz.afterstate_prediction_function.value)
    blob += p + [pol.weight, pol.bias, val.weight, val.bias]
    p, rew = seq(mz.dynamics_function.reward); _, st = seq(mz.dynamics_function.next_state_normalized)
    blob += p + [rew.weight, rew.bias, st.weight, st.bias]
    p, out = seq(mz.encoder_function.encoder); blob += p + [out.weight, out.bias]
    return np.concatenate([t.detach().float().numpy().ravel() for t in blob]).astype(np.float32)


def net_case(name, mz, n_rows, seed, N=50, B=2, K=2):
    """Real-MLP fixture: packed weights, per-function input/output vectors from the reference's
    *_inference methods (muzero_model.py:802-909) and the Encoder forward, plus B recorded searches."""
    g = torch.Generator().manual_seed(seed)
    obs_dim, A, S = mz.observation_dimension, mz.action_dimension, mz.state_dimension
    obs = torch.randn(n_rows, obs_dim, generator=g)
    actions = torch.randint(0, A, (n_rows,), generator=g)
    out = {"weights": pack_reference_weights(mz),
           "dims": np.array([obs_dim, A, A, S, mz.hidden_layer_dimension, mz.number_of_hidden_layer], np.int32),
           "obs": obs.numpy(), "actions": actions.numpy().astype(np.int32)}
    rows = {k: [] for k in ("repr_h", "pred_policy", "pred_value", "adyn_h", "apred_policy", "apred_value",
                            "dyn_h", "dyn_reward", "dpred_policy", "dpred_value", "enc_probs", "enc_code")}
    with torch.no_grad():
        for i in range(n_rows):
            h = mz.representation_function_inference(obs[i:i + 1])
            p, v = mz.prediction_function_inference(h)
            ah = mz.afterstate_dynamics_function_inference(h, int(actions[i]))
            ap, av = mz.afterstate_prediction_function_inference(ah)
            r, dh = mz.dynamics_function_inference(ah, int(actions[i]))
            dp, dv = mz.prediction_function_inference(dh)
            mz.encoder_function.eval()
            c_t, c_e = mz.encoder_function(obs[i:i + 1])
            for k, val in (("repr_h", h[0].numpy()), ("pred_policy", p[0]), ("pred_value", v),
                           ("adyn_h", ah[0].numpy()), ("apred_policy", ap[0]), ("apred_value", av),
                           ("dyn_h", dh[0].numpy()), ("dyn_reward", r), ("dpred_policy", dp[0]),
                           ("dpred_value", dv), ("enc_probs", c_e[0].numpy()),
                           ("enc_code", int(c_t[0].argmax()))):
                rows[k].append(np.asarray(val))
    for k, v in rows.items():
        out[k] = np.stack(v).astype(np.int32 if k == "enc_code" else np.float32)
    np.savez_compressed(os.path.join(GOLDEN_DIR, f"net_{name}.npz"), **out)
    # recorded searches with the real network in the loop
    cfg = dict(pb_c_base=19652, pb_c_init=1.25, discount=0.997, root_dirichlet_alpha=0.25,
               root_exploration_fraction=0.25, num_simulations=N, maxium_action_sample=K,
               number_of_player=1, custom_loop=None)
    runs, obs_rows, root_h, sim_h = [], [], [], []
    for b in range(B):
        o = torch.randn(1, obs_dim, generator=g)
        tape, exp, extra = record_run(cfg, mz, seed=seed + b, train=True, obs=o)
        runs.append((tape, exp))
        obs_rows.append(o[0].numpy())
        root_h.append(extra["root_hidden"][0].numpy())
        sim_h.append(np.stack([h[0].numpy() for h in extra["sim_hidden"]]))
        print("  value_sum/prior/reward python types seen:", extra["types"])
    packed = pack_batch(dict(cfg, action_dim=A, chance_dim=A), True, runs)
    packed["obs"] = np.stack(obs_rows).astype(np.float32)
    packed["root_hidden"] = np.stack(root_h).astype(np.float32)
    packed["sim_hidden"] = np.stack(sim_h).astype(np.float32)
    np.savez_compressed(os.path.join(GOLDEN_DIR, f"tree_{name}.npz"), **packed)
    print(f"net_{name}: rows={n_rows} weights={out['weights'].size}")


def vision_case(name, A=4, S=61, H=126, L=4, n_rows=6, seed=21):
    """Vision (ResNet-v2) fixture: blob in the vision hand-off layout + the reference's own inference outputs
    (muzero_model.py:802-909 with is_RGB) on committed inputs.  BatchNorm statistics / affine parameters are
    randomised (an untrained model has mean 0 / var 1, which would not exercise the eval-mode fold)."""
    import warnings
    sys.path.insert(0, os.path.join(os.path.dirname(GOLDEN_DIR), "..", "stochastic-muzero_b200"))
    from stochastic_muzero_b200.weights import pack_vision_weights
    _, ref_model = ref_shim.load()
    torch.manual_seed(seed)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mz = ref_model.Muzero(model_structure="vision_model", observation_space_dimensions=ref_shim.Box(0, 1, (98, 98, 3)),
                              action_space_dimensions=ref_shim.Discrete(A), state_space_dimensions=S,
                              hidden_layer_dimensions=H, number_of_hidden_layer=L, device="cpu", use_amp=False)
    g = torch.Generator().manual_seed(seed)
    for fn in ("representation", "dynamics", "afterstate_dynamics", "prediction", "afterstate_prediction"):
        for m in getattr(mz, fn + "_function").modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                with torch.no_grad():
                    m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.2)
                    m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
                    m.weight.copy_(torch.rand(m.num_features, generator=g) + 0.5)
                    m.bias.copy_(torch.randn(m.num_features, generator=g) * 0.2)
    blob, shape = pack_vision_weights(mz)
    obs = torch.rand(n_rows, 3, 98, 98, generator=g).half().float()     # stored as float16, so make it exact
    actions = torch.randint(0, A, (n_rows,), generator=g)
    rows = {k: [] for k in ("repr_h", "pred_policy", "pred_value", "adyn_h", "apred_policy", "apred_value", "dyn_h",
                            "dyn_reward", "dpred_policy", "dpred_value")}
    with torch.no_grad():
        for i in range(n_rows):
            h = mz.representation_function_inference(obs[i:i + 1])
            p, v = mz.prediction_function_inference(h)
            ah = mz.afterstate_dynamics_function_inference(h, int(actions[i]))
            ap, av = mz.afterstate_prediction_function_inference(ah)
            r, dh = mz.dynamics_function_inference(ah, int(actions[i]))
            dp, dv = mz.prediction_function_inference(dh)
            for k, val in (("repr_h", h[0].numpy()), ("pred_policy", p[0]), ("pred_value", v), ("adyn_h", ah[0].numpy()),
                           ("apred_policy", ap[0]), ("apred_value", av), ("dyn_h", dh[0].numpy()), ("dyn_reward", r),
                           ("dpred_policy", dp[0]), ("dpred_value", dv)):
                rows[k].append(np.asarray(val))
    out = {k: np.stack(v).astype(np.float32) for k, v in rows.items()}
    out.update(weights=blob, dims=np.array([A, S, H, L], np.int32), obs=obs.numpy().astype(np.float16),
               actions=actions.numpy().astype(np.int32))
    np.savez_compressed(os.path.join(GOLDEN_DIR, f"vision_{name}.npz"), **out)
    print(f"vision_{name}: rows={n_rows} weights={blob.size}")


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(1)
    # --- stub-model tapes: the matrix of SURVEY.md §8c --------------------------------------------
    stub_case("a2c2k2_n50_train", A=2, C=2, K=2, N=50, train=True, B=16)
    stub_case("a2c2k2_n50_eval", A=2, C=2, K=2, N=50, train=False, B=8, seed0=100)
    stub_case("a2c2k2_n1", A=2, C=2, K=2, N=1, train=True, B=4, seed0=200)
    stub_case("a2c2k2_n0", A=2, C=2, K=2, N=0, train=True, B=4, seed0=250)
    stub_case("a2c2k2_n11", A=2, C=2, K=2, N=11, train=True, B=8, seed0=300, discount=0.95)
    stub_case("a4c32k32_n100", A=4, C=32, K=32, N=100, train=True, B=8, seed0=400)
    stub_case("a4c32k4_n50", A=4, C=32, K=4, N=50, train=True, B=8, seed0=500, peaky=2.0)
    stub_case("a4c4k2_n50", A=4, C=4, K=2, N=50, train=True, B=8, seed0=600, peaky=3.0)
    stub_case("a4c4k1_n20", A=4, C=4, K=1, N=20, train=True, B=4, seed0=650)
    stub_case("a3c5k3_n30_p2", A=3, C=5, K=3, N=30, train=True, B=8, seed0=700, players=2)
    stub_case("a2c2k2_n30_loop", A=2, C=2, K=2, N=30, train=True, B=6, seed0=800, players=1,
              custom_loop="1>2>1>3")
    stub_case("a2c2k2_n50_bigval", A=2, C=2, K=2, N=50, train=True, B=8, seed0=900, value_scale=25.0,
              discount=1)
    stub_case("a9c9k9_n40", A=9, C=9, K=9, N=40, train=False, B=4, seed0=950, peaky=0.3, pb_c_base=100,
              pb_c_init=0.5)
    # --- real MLP: random init (cfg-450 shape), a small net, and the shipped 450 checkpoint -------
    net_case("mlp450_seed0", ref_shim.make_muzero(seed=0), n_rows=24, seed=11, N=50, B=4)
    net_case("mlp_small", ref_shim.make_muzero(obs_dim=5, action_dim=3, state_dim=11, hidden_dim=14,
                                               n_hidden=2, seed=1), n_rows=16, seed=12, N=20, B=3, K=3)
    net_case("mlp_l0", ref_shim.make_muzero(obs_dim=4, action_dim=2, state_dim=31, hidden_dim=64,
                                            n_hidden=0, seed=2), n_rows=8, seed=13, N=11, B=2)
    net_case("ckpt450", ref_shim.load_checkpoint(450), n_rows=24, seed=14, N=50, B=4)
    vision_case("a4", A=4, S=61, H=126, L=4, n_rows=6, seed=21)
    vision_case("small", A=3, S=21, H=40, L=1, n_rows=4, seed=22)


if __name__ == "__main__":
    main()
