"""Self-play rows: the vector environment against the CartPole-v1 oracle
(oracle/cartpole_oracle.py, gymnasium's published equations), the batched self-play loop on the GPU with its
trajectories replayed through that oracle, and reanalyse replayed through the search + network oracles."""
import math

import numpy as np
import pytest
import torch

from oracle import cartpole_oracle as CP


# ---- the oracle itself: hand-computed known answers -----------------------------------------------------------
def test_cartpole_oracle_known_answers():
    # from rest, push right: temp = 10/1.1, thetaacc = -temp / (0.5 * (4/3 - 0.1/1.1)) = -14.634146..,
    # xacc = temp + 0.05 * 14.634146 / 1.1 = 9.756097..; one Euler step of 0.02 s
    nxt, rew, term = CP.step(np.zeros((1, 4)), np.array([1]))
    np.testing.assert_allclose(nxt[0], [0.0, 0.1951219512195122, 0.0, -0.2926829268292683], rtol=1e-15)
    assert rew[0] == 1.0 and not term[0]
    nxt, _, _ = CP.step(np.zeros((1, 4)), np.array([0]))
    np.testing.assert_allclose(nxt[0], [0.0, -0.1951219512195122, 0.0, 0.2926829268292683], rtol=1e-15)
    # tilted, moving pole (values computed independently with the scalar formulas of cartpole.py)
    s = np.array([[0.01, -0.02, 0.03, 0.04]])
    np.testing.assert_allclose(CP.step(s, [1])[0][0], [0.0096, 0.17467919574755525, 0.0308, -0.2430687179600081], rtol=1e-13)
    np.testing.assert_allclose(CP.step(s, [0])[0][0], [0.0096, -0.21553901710278936, 0.0308, 0.34199522377603914], rtol=1e-13)
    # leaving the track / the 12 degree cone terminates; the step that crosses the threshold is the terminal one
    nxt, _, term = CP.step(np.array([[2.39, 1.0, -0.2, -1.5]]), [0])
    np.testing.assert_allclose(nxt[0], [2.41, 0.807789466233749, -0.23, -1.275840103173756], rtol=1e-13)
    assert term[0]
    assert CP.step(np.array([[0.0, 0.0, CP.THETA_THRESHOLD - 1e-9, 1.0]]), [1])[2][0]
    assert not CP.step(np.array([[0.0, 0.0, 0.1, 0.0]]), [1])[2][0]
    # alternating pushes keep a centred pole up for a while; the TimeLimit of CartPole-v1 ends an episode at 500 steps
    traj, done_at = CP.rollout(np.zeros(4), [1, 0] * 300)
    assert done_at is not None and done_at < 500 and len(traj) == done_at + 2


def _check_against_oracle(env, steps, seed):
    """Every device step equals one oracle step from the same state (float32 vs float64 arithmetic), the done mask
    is the oracle's termination or the 500-step limit, finished environments restart inside +-0.05."""
    g = np.random.default_rng(seed)
    n = env.n
    state = env.state.double().cpu().numpy()
    for _ in range(steps):
        a = g.integers(0, 2, n)
        prev_steps = env.steps.cpu().numpy().copy()
        obs, rew, done = env.step(torch.from_numpy(a).to(env.device))
        exp, exp_rew, exp_term = CP.step(state, a)
        np.testing.assert_allclose(obs.double().cpu().numpy(), exp, rtol=2e-5, atol=2e-6)
        assert (rew.cpu().numpy() == exp_rew).all()
        # thresholds are compared in float32 on the device: allow disagreement only within rounding of the threshold
        margin = np.minimum(np.abs(np.abs(exp[:, 0]) - CP.X_THRESHOLD), np.abs(np.abs(exp[:, 2]) - CP.THETA_THRESHOLD))
        exp_done = exp_term | (prev_steps + 1 >= CP.MAX_EPISODE_STEPS)
        d = done.cpu().numpy()
        assert (d == exp_done)[margin > 1e-5].all()
        obs = env.reset(done)
        state = obs.double().cpu().numpy()
        assert (np.abs(state[d]) <= 0.05 + 1e-7).all() and (env.steps[done] == 0).all()
        assert (env.steps.cpu().numpy()[~d] == prev_steps[~d] + 1).all()


def test_vector_cartpole_matches_the_oracle_step_by_step():
    from stochastic_muzero_b200.selfplay import VectorCartPole
    _check_against_oracle(VectorCartPole(64, device="cpu", seed=3), steps=120, seed=0)


def test_vector_cartpole_time_limit():
    from stochastic_muzero_b200.selfplay import VectorCartPole
    env = VectorCartPole(2, device="cpu", seed=1)
    env.steps[:] = CP.MAX_EPISODE_STEPS - 1
    _, _, done = env.step(torch.tensor([0, 1]))
    assert done.all()                                    # truncated by the 500-step limit of CartPole-v1


@pytest.mark.gpu
def test_vector_cartpole_on_device_matches_the_oracle():
    from stochastic_muzero_b200.selfplay import VectorCartPole
    _check_against_oracle(VectorCartPole(2048, device="cuda", seed=5), steps=200, seed=1)


@pytest.mark.gpu
def test_batched_selfplay_loop_on_device():
    from stochastic_muzero_b200 import ModelShape, Monte_carlo_tree_search, PackedModel, random_blob
    from stochastic_muzero_b200.selfplay import SelfPlay, VectorCartPole
    shape = ModelShape(4, 2, 2, 61, 126, 4)
    mcts = Monte_carlo_tree_search(discount=0.997, num_simulations=16, maxium_action_sample=2, net="bf16", seed=1)
    env = VectorCartPole(512, device="cuda", seed=0)
    T = 40
    out = SelfPlay(mcts, PackedModel(random_blob(shape, 0), shape), env, horizon=T).run(temperature=1.0)
    assert out["actions"].shape == (T, 512) and set(out["actions"].unique().tolist()) <= {0, 1}
    assert torch.allclose(out["child_visits"].sum(2), torch.ones(T, 512, dtype=torch.float64, device="cuda"))
    assert torch.isfinite(out["root_values"]).all() and (out["rewards"] == 1).all()
    # with a random policy the pole falls within ~10-40 steps: games must have ended and restarted
    ends = out["dones"].sum().item()
    assert ends > 256, f"only {ends} episode ends in {T} moves of 512 random-policy games"
    # the stored trajectories ARE CartPole-v1 trajectories: replay every transition through the oracle
    obs = out["observations"].double().cpu().numpy()
    act, dones = out["actions"].cpu().numpy(), out["dones"].cpu().numpy()
    for t in range(T - 1):
        exp, _, term = CP.step(obs[t], act[t])
        cont = ~dones[t]
        np.testing.assert_allclose(obs[t + 1][cont], exp[cont], rtol=2e-5, atol=2e-6)
        assert not term[cont & (np.abs(np.abs(exp[:, 2]) - CP.THETA_THRESHOLD) > 1e-5) &
                        (np.abs(np.abs(exp[:, 0]) - CP.X_THRESHOLD) > 1e-5)].any(), "a terminated game was continued"
        assert (np.abs(obs[t + 1][dones[t]]) <= 0.05 + 1e-6).all()           # finished games restart from a fresh state
        ended = dones[t] & (np.abs(np.abs(exp[:, 2]) - CP.THETA_THRESHOLD) > 1e-5)
        assert term[ended].all(), "a game was reset although the oracle says it goes on"     # (no 500-step limit in 40 moves)


@pytest.mark.gpu
def test_reanalyse_targets_equal_an_oracle_replay_of_the_stored_observations():
    """Reanalyse (self_play.py:30-44, game.py:112-116, replay_buffer.py:229-266): stored observations searched again
    with the current weights.  The targets must be what the reference's search gives for each stored observation:
    every sampled position is replayed on the CPU by the search oracle driving the network oracle, with the Philox
    stream the engine used for that position (seed of its chunk, tree id = index inside the chunk)."""
    import golden_io
    from oracle import mcts_oracle as O
    from oracle import net_oracle as NO
    from stochastic_muzero_b200 import ModelShape, Monte_carlo_tree_search, PackedModel
    from stochastic_muzero_b200.selfplay import reanalyse
    zn = golden_io.load_net_case("ckpt450")
    dims = [int(v) for v in zn["dims"]]
    model = PackedModel(zn["weights"], ModelShape(*dims))
    T, B, N, chunk, seed = 6, 50, 12, 128, 2
    kw = dict(pb_c_base=19652, pb_c_init=1.25, discount=0.997, root_dirichlet_alpha=0.25,
              root_exploration_fraction=0.25, num_simulations=N, maxium_action_sample=2)
    mcts = Monte_carlo_tree_search(**kw, net="fp32", seed=seed, tree_id_offset=0)
    obs = torch.randn(T, B, 4, generator=torch.Generator().manual_seed(7)) * 0.1
    out = reanalyse(mcts, model, obs, chunk=chunk, train=False)            # no Dirichlet noise: fully reproducible
    assert out["child_visits"].shape == (T, B, 2) and out["root_values"].shape == (T, B)
    assert torch.allclose(out["child_visits"].sum(2), torch.ones(T, B, dtype=torch.float64, device="cuda"))
    visits = out["child_visits"].reshape(-1, 2).cpu().numpy()
    values = out["root_values"].reshape(-1).cpu().numpy()
    net = NO.NetOracle(zn["weights"], *dims)
    cfg = O.SearchConfig(**kw)
    flat = obs.reshape(-1, 4).numpy()
    same, sample = 0, list(range(0, T * B, 7))
    for p in sample:
        run = p // chunk + 1                                              # reanalyse's k-th run_batch call
        run_seed = (seed + 0x9E3779B97F4A7C15 * run) & 0xFFFFFFFFFFFFFFFF  # Monte_carlo_tree_search._next_seed
        tree = O.search(cfg, NO.NetModel(net, flat[p]), O.PhiloxUniforms(run_seed, p % chunk), train=False)
        kids = [n for n in range(len(tree.visit)) if tree.depth[n] == 1]
        v = np.array([tree.visit[n] for n in kids], np.float64)
        if np.array_equal(v / v.sum(), visits[p]):
            same += 1
            root_value = np.float32(tree.value_sum[0]) / np.float32(tree.visit[0])
            np.testing.assert_allclose(values[p], root_value, rtol=5e-5, atol=1e-5)
    # numpy fp32 network vs CUDA fp32 network: a score tie below 1e-6 may flip a visit in a few positions
    assert same >= 0.9 * len(sample), f"only {same}/{len(sample)} reanalysed positions match the oracle replay"
