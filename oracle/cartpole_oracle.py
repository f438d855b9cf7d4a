"""TEST INFRASTRUCTURE — CPU oracle for the vector environment of the batched self-play row.

The reference steps gymnasium environments (game.py:96-131: ``env.reset(seed=...)`` / ``env.step(action)``; its
CartPole experiments use ``CartPole-v1``).  gymnasium is a third-party dependency pinned at ``gymnasium[all]==0.27.0``
(/root/reference/requirements.txt:16) and absent from this image, so the published algorithm of
``gymnasium/envs/classic_control/cartpole.py`` (0.27.0; Barto, Sutton & Anderson 1983) is restated here in
float64, exactly as that file computes it:

    force      = +10 N for action 1, -10 N for action 0
    temp       = (force + polemass_length * theta_dot**2 * sin(theta)) / total_mass
    thetaacc   = (gravity * sin(theta) - cos(theta) * temp)
                 / (length * (4/3 - masspole * cos(theta)**2 / total_mass))
    xacc       = temp - polemass_length * thetaacc * cos(theta) / total_mass
    euler:  x += tau * x_dot;  x_dot += tau * xacc;  theta += tau * theta_dot;  theta_dot += tau * thetaacc
    terminated = |x| > 2.4  or  |theta| > 12 * 2 * pi / 360;   reward = 1.0 per step
    CartPole-v1 registration: max_episode_steps = 500 (TimeLimit wrapper -> truncated)
    reset: the four state variables ~ U(-0.05, 0.05)

with gravity 9.8, masscart 1.0, masspole 0.1, total_mass 1.1, length 0.5 (half the pole), polemass_length 0.05,
force_mag 10.0, tau 0.02.  "Parity unpinned" by anything reference-owned (the reference holds no environment
fixtures); pinned here by hand-computed known answers (tests/test_selfplay.py).
"""
import math

import numpy as np

GRAVITY, MASSCART, MASSPOLE, LENGTH, FORCE_MAG, TAU = 9.8, 1.0, 0.1, 0.5, 10.0, 0.02
TOTAL_MASS = MASSPOLE + MASSCART
POLEMASS_LENGTH = MASSPOLE * LENGTH
X_THRESHOLD = 2.4
THETA_THRESHOLD = 12 * 2 * math.pi / 360
MAX_EPISODE_STEPS = 500


def step(state, action):
    """One Euler step for a batch: state float64 [n, 4], action int [n] -> (next_state, reward, terminated)."""
    state = np.asarray(state, dtype=np.float64)
    x, x_dot, theta, theta_dot = state.T
    force = np.where(np.asarray(action) == 1, FORCE_MAG, -FORCE_MAG)
    costheta, sintheta = np.cos(theta), np.sin(theta)
    temp = (force + POLEMASS_LENGTH * np.square(theta_dot) * sintheta) / TOTAL_MASS
    thetaacc = (GRAVITY * sintheta - costheta * temp) / (LENGTH * (4.0 / 3.0 - MASSPOLE * np.square(costheta) / TOTAL_MASS))
    xacc = temp - POLEMASS_LENGTH * thetaacc * costheta / TOTAL_MASS
    nxt = np.stack([x + TAU * x_dot, x_dot + TAU * xacc, theta + TAU * theta_dot, theta_dot + TAU * thetaacc], axis=1)
    terminated = (nxt[:, 0] < -X_THRESHOLD) | (nxt[:, 0] > X_THRESHOLD) | \
                 (nxt[:, 2] < -THETA_THRESHOLD) | (nxt[:, 2] > THETA_THRESHOLD)
    return nxt, np.ones(len(nxt)), terminated


def rollout(state, actions):
    """Follow one environment: state [4], actions [T] -> states [T+1, 4], done step (or None).  done = terminated or
    the 500-step TimeLimit of CartPole-v1."""
    traj, done_at = [np.asarray(state, np.float64)], None
    for t, a in enumerate(actions):
        nxt, _, term = step(traj[-1][None], np.array([a]))
        traj.append(nxt[0])
        if term[0] or t + 1 >= MAX_EPISODE_STEPS:
            done_at = t
            break
    return np.stack(traj), done_at
