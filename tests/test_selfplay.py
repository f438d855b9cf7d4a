"""Self-play row: vector environment equations on CPU, batched self-play loop on the GPU."""
import math

import numpy as np
import pytest
import torch


def _cartpole_numpy(state, action):
    """CartPole-v1 equations of motion, scalar numpy restatement (Barto, Sutton & Anderson 1983)."""
    g, mc, mp, l, f, tau = 9.8, 1.0, 0.1, 0.5, 10.0, 0.02
    x, xd, th, thd = state
    force = f if action == 1 else -f
    tm, pml = mc + mp, mp * l
    temp = (force + pml * thd ** 2 * math.sin(th)) / tm
    tha = (g * math.sin(th) - math.cos(th) * temp) / (l * (4.0 / 3.0 - mp * math.cos(th) ** 2 / tm))
    xa = temp - pml * tha * math.cos(th) / tm
    return np.array([x + tau * xd, xd + tau * xa, th + tau * thd, thd + tau * tha])


def test_vector_cartpole_follows_the_scalar_equations():
    from stochastic_muzero_b200.selfplay import VectorCartPole
    env = VectorCartPole(64, device="cpu", seed=3)
    state = env.state.clone().double().numpy()
    g = np.random.default_rng(0)
    for _ in range(30):
        a = g.integers(0, 2, 64)
        obs, rew, done = env.step(torch.from_numpy(a))
        state = np.stack([_cartpole_numpy(s, int(k)) for s, k in zip(state, a)])
        np.testing.assert_allclose(obs.numpy(), state, rtol=2e-4, atol=2e-5)
        assert (rew == 1).all()
        exp_done = (np.abs(state[:, 0]) > 2.4) | (np.abs(state[:, 2]) > 12 * 2 * math.pi / 360)
        assert np.array_equal(done.numpy(), exp_done)
        obs = env.reset(done)
        state = obs.double().numpy()
        assert (np.abs(state[exp_done]) <= 0.05).all() and (env.steps[done] == 0).all()


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
    # after a done, the next observation is a fresh start (|state| <= 0.05)
    d = out["dones"][:-1]
    nxt = out["observations"][1:][d]
    assert (nxt.abs() <= 0.05 + 1e-6).all()


@pytest.mark.gpu
def test_reanalyse_recomputes_targets_for_stored_positions():
    from stochastic_muzero_b200 import ModelShape, Monte_carlo_tree_search, PackedModel, random_blob
    from stochastic_muzero_b200.selfplay import reanalyse
    shape = ModelShape(4, 2, 2, 61, 126, 4)
    model = PackedModel(random_blob(shape, 0), shape)
    mcts = Monte_carlo_tree_search(discount=0.997, num_simulations=10, maxium_action_sample=2, net="fp32", seed=2)
    obs = torch.randn(6, 50, 4)
    out = reanalyse(mcts, model, obs, chunk=128)
    assert out["child_visits"].shape == (6, 50, 2) and out["root_values"].shape == (6, 50)
    assert torch.allclose(out["child_visits"].sum(2), torch.ones(6, 50, dtype=torch.float64, device="cuda"))
    # position (t, b) is searched as an independent tree: same answer as searching it alone with the same draws
    mcts2 = Monte_carlo_tree_search(discount=0.997, num_simulations=10, maxium_action_sample=2, net="fp32", seed=2)
    first = mcts2.run_batch(obs.reshape(300, 4)[:128], model, train=True)
    assert torch.equal(first.select_actions(0.0)["stored_policy"].reshape(-1, 2), out["child_visits"].reshape(-1, 2)[:128])
