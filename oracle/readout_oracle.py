"""TEST INFRASTRUCTURE — CPU oracle for the step after the search: root read-out and action selection.
NOT PRODUCT CODE.

numpy restatement of /root/reference/game.py:179-216 (`store_search_statistics`,
`policy_action_reward_from_tree`, `softmax_stable`, `select_action`), float64 throughout like the
reference.  Pinned against the reference's own methods by the `readout_*` arrays oracle/make_golden.py
stores in tests/golden/tree_*.npz.
"""
import numpy as np


def stored_policy(visits, priors):
    """game.py:179-191: the policy appended to Game.child_visits (no temperature)."""
    v = np.asarray(visits, dtype=np.float64)
    if v.sum() >= 3:
        return v / v.sum()
    p = np.asarray(priors, dtype=np.float64)
    return p / p.sum()                      # softmax_stable(policy, temperature=0): temperature < 0.3 => no power


def step_policy(visits, priors, temperature):
    """game.py:198-211 + :226-232: visit counts (priors if sum <= 1), ** (1/T) only if T >= 0.3, normalised."""
    p = np.asarray(visits, dtype=np.float64)
    if p.sum() <= 1:
        p = np.asarray(priors, dtype=np.float64)
    if temperature >= 0.3:
        p = p ** (1 / temperature)
    return p / p.sum()


def select_action(policy, temperature, u):
    """game.py:211-216: sample (np.random.choice == searchsorted on the normalised cumsum, one uniform u) when
    T > 0.1 or all entries are equal, else first argmax.  Returns (index into the children, sampled?)."""
    if temperature > 0.1 or len(set(policy)) == 1:
        cdf = np.cumsum(policy)
        cdf /= cdf[-1]
        return int(np.searchsorted(cdf, u, side="right")), True
    return int(np.argmax(policy)), False
