"""TEST INFRASTRUCTURE — CPU oracle for the Stochastic-MuZero search path.  NOT PRODUCT CODE.

A numpy restatement of the reference's ``Monte_carlo_tree_search.run``
(/root/reference/monte_carlo_tree_search.py:311-349) written against an explicit random-number tape
and an explicit network-output source, so that the CUDA engine and the reference can be compared on
*identical recorded inputs*.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module; the product package never does.

Pinning status: the reference ships no tests or golden vectors for this path, so the
oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, run in the build container by
``oracle/make_golden.py`` (record/replay of ``np.random`` draws and network outputs) and committed
under ``tests/golden/``; ``tests/test_oracle_golden.py`` replays every committed tape through this
module and demands bit-exact tree statistics.

dtype semantics are the ones observed with numpy >= 2 (NEP 50 weak python scalars):
tree values are float32, pUCT prior term float64.  Every cast is written out explicitly so that the
restatement does not depend on the promotion rules of the numpy that happens to be installed.
"""
from __future__ import annotations

import dataclasses
from typing import List, Optional, Sequence

import numpy as np

f32 = np.float32
f64 = np.float64


# --------------------------------------------------------------------------------------------------
# configuration (ctor arguments of the reference, monte_carlo_tree_search.py:76-85)
# --------------------------------------------------------------------------------------------------
@dataclasses.dataclass
class SearchConfig:
    pb_c_base: int = 19652
    pb_c_init: float = 1.25
    discount: float = 0.95
    root_dirichlet_alpha: float = 0.25
    root_exploration_fraction: float = 0.25
    num_simulations: int = 10
    maxium_action_sample: int = 2
    number_of_player: int = 1
    custom_loop: Optional[str] = None

    def cycle_map(self) -> np.ndarray:
        """Player_cycle.{modular,custom}_cycle (monte_carlo_tree_search.py:50-58)."""
        if self.custom_loop is not None:
            return np.array([float(i) for i in self.custom_loop.split(">")], dtype=np.float32)
        return np.arange(self.number_of_player, dtype=np.int64)


# --------------------------------------------------------------------------------------------------
# random sources: every draw of the search reduces to RandomState.random_sample() doubles (T5)
# --------------------------------------------------------------------------------------------------
class TapeUniforms:
    """Replays a recorded stream of uniform doubles in consumption order."""

    def __init__(self, values: Sequence[float]):
        self.values = np.asarray(values, dtype=np.float64)
        self.cursor = 0

    def next(self) -> np.float64:
        v = self.values[self.cursor]
        self.cursor += 1
        return v


class MTUniforms:
    """Live MT19937 stream (numpy legacy RandomState), the generator the reference draws from."""

    def __init__(self, seed: int):
        self.rs = np.random.RandomState(seed)
        self.log: List[float] = []

    def next(self) -> np.float64:
        v = np.float64(self.rs.random_sample())
        self.log.append(float(v))
        return v


_PHILOX_M0, _PHILOX_M1 = 0xD2511F53, 0xCD9E8D57
_PHILOX_W0, _PHILOX_W1 = 0x9E3779B9, 0xBB67AE85


def philox4x32_10(counter, key):
    """Philox4x32-10 (Salmon et al. 2011), same constants as the CUDA engine's device RNG."""
    c0, c1, c2, c3 = counter
    k0, k1 = key
    for _ in range(10):
        p0 = _PHILOX_M0 * c0
        p1 = _PHILOX_M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, \
                         ((p0 >> 32) ^ c3 ^ k1) & 0xFFFFFFFF, p0 & 0xFFFFFFFF
        k0 = (k0 + _PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + _PHILOX_W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def philox_uniform(seed: int, tree_id: int, index: int, stream: int = 0) -> np.float64:
    """The engine's production uniform #index of tree #tree_id: 53-bit double in [0,1).

    counter = (index>>1, stream, tree_id_lo, tree_id_hi), key = (seed_lo, seed_hi); a Philox block
    yields two doubles, each built like numpy's random_sample: (a>>5)*2^26 + (b>>6), over 2^53.
    """
    c = ((index >> 1) & 0xFFFFFFFF, stream & 0xFFFFFFFF, tree_id & 0xFFFFFFFF, (tree_id >> 32) & 0xFFFFFFFF)
    k = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    r = philox4x32_10(c, k)
    a, b = (r[0], r[1]) if (index & 1) == 0 else (r[2], r[3])
    return np.float64(((a >> 5) * 67108864 + (b >> 6)) / 9007199254740992.0)


class PhiloxUniforms:
    def __init__(self, seed: int, tree_id: int):
        self.seed, self.tree_id, self.cursor = seed, tree_id, 0

    def next(self) -> np.float64:
        v = philox_uniform(self.seed, self.tree_id, self.cursor)
        self.cursor += 1
        return v


# --------------------------------------------------------------------------------------------------
# network-output sources
# --------------------------------------------------------------------------------------------------
class TapeModel:
    """Recorded network outputs, indexed by simulation (identical draws => identical visiting order)."""

    def __init__(self, root_policy, sim_policy, sim_width, sim_value, sim_reward):
        self.root_policy = np.asarray(root_policy, dtype=np.float32)
        self.sim_policy = np.asarray(sim_policy, dtype=np.float32)
        self.sim_width = np.asarray(sim_width, dtype=np.int64)
        self.sim_value = np.asarray(sim_value, dtype=np.float32)
        self.sim_reward = np.asarray(sim_reward, dtype=np.float32)

    def root(self):
        return None, self.root_policy, f32(0)

    def afterstate(self, sim, parent_hidden, action):
        w = int(self.sim_width[sim])
        return None, self.sim_policy[sim, :w], self.sim_value[sim]

    def dynamics(self, sim, parent_hidden, code):
        w = int(self.sim_width[sim])
        return None, self.sim_policy[sim, :w], self.sim_value[sim], self.sim_reward[sim]


# --------------------------------------------------------------------------------------------------
# numpy primitives whose rounding order matters
# --------------------------------------------------------------------------------------------------
def normalise_policy(policy: np.ndarray) -> np.ndarray:
    """(policy + 1e-12) / sum — float32 throughout (monte_carlo_tree_search.py:205-206, :291-292)."""
    p = (policy.astype(np.float32) + f32(1e-12)).astype(np.float32)
    return (p / p.sum(dtype=np.float32)).astype(np.float32)


def choice_cdf(p32: np.ndarray) -> np.ndarray:
    """cdf used by RandomState.choice: float64 cumsum of p, divided by its last entry."""
    cdf = np.cumsum(p32.astype(np.float64))
    cdf /= cdf[-1]
    return cdf


def choice_without_replacement(n: int, size: int, p32: np.ndarray, rng) -> List[int]:
    """np.random.choice(n, size, p=p, replace=False) as numpy's legacy RandomState implements it:
    rounds of (size - found) uniforms, searchsorted(side='right') on the renormalised cdf with found
    entries zeroed, first occurrences kept in draw order.  Returns indices in FOUND order."""
    p = p32.astype(np.float64).copy()
    found: List[int] = []
    while len(found) < size:
        m = size - len(found)
        xs = [rng.next() for _ in range(m)]
        if found:
            p[found] = 0.0
        cdf = np.cumsum(p)
        cdf /= cdf[-1]
        for x in xs:
            idx = int(np.searchsorted(cdf, x, side="right"))
            if idx not in found:
                found.append(idx)
    return found


def smoothed_chance_probs(priors32: np.ndarray) -> np.ndarray:
    """select_child's chance branch (monte_carlo_tree_search.py:251-253), float32."""
    probs = priors32.astype(np.float32)
    one_minus = (f32(1) - probs).astype(np.float32) + f32(1e-12)
    remainder = np.abs(one_minus.astype(np.float32).mean(dtype=np.float32))
    shifted = (probs + f32(remainder)).astype(np.float32)
    return (shifted / shifted.sum(dtype=np.float32)).astype(np.float32)


# --------------------------------------------------------------------------------------------------
# the tree
# --------------------------------------------------------------------------------------------------
class Tree:
    """Flat node table; node 0 is the root.  Mirrors Node (monte_carlo_tree_search.py:6-21)."""

    def __init__(self):
        self.visit: List[int] = []
        self.value_sum: List[np.float32] = []
        self.reward: List[np.float32] = []
        self.prior: List[object] = []          # float32, or float64 for noised root children
        self.key: List[int] = []
        self.depth: List[int] = []
        self.is_chance: List[bool] = []
        self.to_play: List[int] = []
        self.children: List[List[int]] = []    # ascending key order
        self.hidden: List[object] = []
        self.vmin = float("inf")
        self.vmax = -float("inf")
        self.paths: List[List[int]] = []       # per simulation: keys chosen root->leaf

    def add(self, prior, key, depth, is_chance, to_play) -> int:
        self.visit.append(0)
        self.value_sum.append(f32(0))
        self.reward.append(f32(0))
        self.prior.append(prior)
        self.key.append(int(key))
        self.depth.append(depth)
        self.is_chance.append(bool(is_chance))
        self.to_play.append(int(to_play))
        self.children.append([])
        self.hidden.append(None)
        return len(self.visit) - 1

    def value(self, n: int) -> np.float32:
        return f32(0) if self.visit[n] == 0 else f32(self.value_sum[n] / f32(self.visit[n]))

    # canonical dump: depth-first, children in ascending key order
    def dump(self) -> dict:
        order: List[int] = []
        stack = [0]
        while stack:
            n = stack.pop()
            order.append(n)
            stack.extend(reversed(self.children[n]))
        return {
            "depth": np.array([self.depth[n] for n in order], dtype=np.int32),
            "key": np.array([self.key[n] for n in order], dtype=np.int32),
            "visit": np.array([self.visit[n] for n in order], dtype=np.int32),
            "value_sum": np.array([self.value_sum[n] for n in order], dtype=np.float32),
            "reward": np.array([self.reward[n] for n in order], dtype=np.float32),
            "prior": np.array([np.float64(self.prior[n]) for n in order], dtype=np.float64),
            "is_chance": np.array([self.is_chance[n] for n in order], dtype=np.int8),
            "to_play": np.array([self.to_play[n] for n in order], dtype=np.int32),
            "expanded": np.array([len(self.children[n]) > 0 for n in order], dtype=np.int8),
            "minmax": np.array([self.vmin, self.vmax], dtype=np.float32),
        }


def ucb_score(cfg: SearchConfig, tree: Tree, parent: int, child: int, u: np.float64) -> np.float64:
    """monte_carlo_tree_search.py:235-243 with the dtypes of SURVEY.md §8a M7/T3/T4."""
    n_parent, n_child = tree.visit[parent], tree.visit[child]
    pb_c = np.log((n_parent + cfg.pb_c_base + 1) / cfg.pb_c_base) + cfg.pb_c_init            # f64
    prior_score = (np.sqrt(f64(n_parent)) * pb_c * f64(tree.prior[child])) / f64(n_child + 1)  # f64
    if n_child > 0:
        v = f32(tree.reward[child] + f32(f32(cfg.discount) * tree.value(child)))               # f32, 2 roundings
        if tree.vmax > tree.vmin:
            v = f32(f32(v - f32(tree.vmin)) / f32(f32(tree.vmax) - f32(tree.vmin)))
        value_score = f64(v)
    else:
        value_score = f64(0)
    noise = f64(1e-7) + f64(2e-7 - 1e-7) * u      # RandomState.uniform: low + (high-low)*random_sample()
    return f64(f64(prior_score + value_score) + noise)


def search(cfg: SearchConfig, model, rng, train: bool = True, dirichlet: Optional[np.ndarray] = None,
           root_to_play: int = 0, dirichlet_fn=None) -> Tree:
    """One Monte_carlo_tree_search.run (monte_carlo_tree_search.py:311-349).

    model : object with root() / afterstate(sim, h, a) / dynamics(sim, h, c)
    rng   : object with next() -> uniform double in [0,1)
    dirichlet : recorded np.random.dirichlet output (float64 [A]); or dirichlet_fn(A) to draw one.
    """
    cycle = cfg.cycle_map()
    n_cycle = len(cycle)

    def next_player(p):      # Player_cycle.proximate_player_step (:60-61)
        return (p + 1) % n_cycle

    def in_play(p):          # Player_cycle.player_in_play (:71-72)
        return cycle[p % n_cycle]

    tree = Tree()
    root = tree.add(prior=0, key=-1, depth=0, is_chance=False, to_play=root_to_play)
    hidden, policy, _root_value = model.root()          # root value is discarded (T7)
    tree.hidden[root] = hidden

    # expand_the_children_of_the_root_node (:203-211): all A actions, ascending; draws are consumed
    p = normalise_policy(policy)
    n = p.shape[0]
    for i in sorted(choice_without_replacement(n, n, p, rng)):
        c = tree.add(prior=p[i], key=i, depth=1, is_chance=False, to_play=next_player(root_to_play))
        tree.children[root].append(c)

    # add_exploration_noise_at_the_root (:214-225)
    if cfg.num_simulations == 0:
        train = False
    if train:
        noise = np.asarray(dirichlet if dirichlet is not None else dirichlet_fn(n), dtype=np.float64)
        frac = cfg.root_exploration_fraction
        for c, nz in zip(tree.children[root], noise):
            scaled = f32(f32(tree.prior[c]) * f32(1 - frac))          # f32 * weak python float
            tree.prior[c] = f64(f64(scaled) + f64(nz) * f64(frac))      # + float64 => float64 prior

    for sim in range(cfg.num_simulations):
        node = root
        path = [root]
        keys: List[int] = []
        # choice_node_to_expand_using_max_ucb_score (:262-267)
        while tree.children[node]:
            kids = tree.children[node]
            if tree.is_chance[node]:
                probs = smoothed_chance_probs(np.array([tree.prior[c] for c in kids], dtype=np.float32))
                cdf = choice_cdf(probs)
                pick = kids[int(np.searchsorted(cdf, rng.next(), side="right"))]
            else:
                best = None
                for c in kids:                       # ascending key; one fresh uniform per child
                    s = ucb_score(cfg, tree, node, c, rng.next())
                    if best is None or (s, tree.key[c]) > best[:2]:   # max() of (score, action, ...) (T6)
                        best = (s, tree.key[c], c)
                pick = best[2]
            node = pick
            keys.append(tree.key[pick])
            path.append(pick)
        tree.paths.append(keys)
        parent = path[-2]

        # leaf evaluation (:333-342)
        if tree.is_chance[parent]:
            h, policy, value, reward = model.dynamics(sim, tree.hidden[parent], keys[-1])
            tree.reward[node] = f32(reward)
            child_is_chance = False
        else:
            h, policy, value = model.afterstate(sim, tree.hidden[parent], keys[-1])
            child_is_chance = True
        tree.hidden[node] = h

        # create_new_node_in_the_chosen_node_with_action_and_policy (:289-297)
        p = normalise_policy(policy)
        n = p.shape[0]
        bound = min(cfg.maxium_action_sample, n)
        for i in sorted(choice_without_replacement(n, bound, p, rng)):
            tp = tree.to_play[node] if child_is_chance else next_player(tree.to_play[node])
            c = tree.add(prior=p[i], key=i, depth=tree.depth[node] + 1, is_chance=child_is_chance, to_play=tp)
            tree.children[node].append(c)

        # back_propagate_and_update_min_max_bound (:299-308)
        value = f32(value)
        for b in reversed(path):
            same = in_play(tree.to_play[root]) == in_play(tree.to_play[b])
            tree.value_sum[b] = f32(tree.value_sum[b] + (value if same else f32(-value)))
            tree.visit[b] += 1
            nv = tree.value(b)
            tree.vmax = max(tree.vmax, nv)
            tree.vmin = min(tree.vmin, nv)
            value = f32(tree.reward[b] + f32(f32(cfg.discount) * value))
    return tree
