"""CPU-side checks: the C-ABI library loads and exports every symbol include/smz.h declares, the
weight hand-off layout, Player_cycle tables and the drop-in call surface."""
import os
import re

import numpy as np
import pytest

import golden_io
from fake_muzero import FakeMuzero
from oracle import mcts_oracle as O
from oracle import net_oracle as NO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from stochastic_muzero_b200 import _lib
    header = open(os.path.join(ROOT, "include", "smz.h")).read()
    declared = set(re.findall(r"\b(smz_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found in include/smz.h"
    lib = _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"libsmz.so does not export {name}"
    assert declared == set(_lib.SIGNATURES), "ctypes SIGNATURES and include/smz.h disagree"


def test_create_rejects_bad_config_without_gpu_work():
    import ctypes as C
    from stochastic_muzero_b200 import _lib
    lib = _lib.load()
    cfg = _lib.smz_config()
    cfg.abi_version = 999
    h = C.c_void_p()
    assert lib.smz_create(C.byref(cfg), C.byref(h)) == _lib.SMZ_E_INVALID_ARG
    assert b"abi_version" in lib.smz_last_error()
    cfg.abi_version = _lib.SMZ_ABI_VERSION
    cfg.max_trees, cfg.num_simulations, cfg.action_dim, cfg.chance_dim = 4, 5, 64, 2
    cfg.max_action_sample, cfg.pb_c_base, cfg.pb_c_init, cfg.discount = 2, 19652, 1.25, 0.99
    assert lib.smz_create(C.byref(cfg), C.byref(h)) == _lib.SMZ_E_CAPACITY


def test_blob_layout_matches_oracle_layout():
    from stochastic_muzero_b200.weights import ModelShape, blob_layout
    for dims in [(4, 2, 2, 61, 126, 4), (16, 4, 32, 61, 126, 4), (5, 3, 3, 11, 14, 2), (4, 2, 2, 31, 64, 0)]:
        mine, total = blob_layout(ModelShape(*dims))
        theirs, total2 = NO.blob_layout(*dims)
        assert total == total2 and mine == theirs


@pytest.mark.parametrize("name", golden_io.net_cases())
def test_pack_weights_roundtrip(name):
    from stochastic_muzero_b200.weights import pack_weights, shape_of
    z = golden_io.load_net_case(name)
    dims = [int(v) for v in z["dims"]]
    fake = FakeMuzero(z["weights"], *dims)
    blob, shape = pack_weights(fake)
    assert (shape.obs_dim, shape.action_dim, shape.chance_dim, shape.state_dim, shape.hidden_dim,
            shape.num_hidden_layers) == tuple(dims)
    assert np.array_equal(blob, z["weights"])
    assert shape_of(fake) == shape


@pytest.mark.parametrize("players,loop", [(1, None), (2, None), (3, None), (1, "1>2>1>3")])
def test_player_tables_follow_oracle_tree(players, loop):
    """to_play / sign per depth must equal what the oracle's explicit Player_cycle bookkeeping gives."""
    from stochastic_muzero_b200.engine import player_tables
    N = 24
    sign, to_play = player_tables(players, loop, N + 2)
    cfg = O.SearchConfig(num_simulations=N, maxium_action_sample=2, number_of_player=players, custom_loop=loop,
                         discount=0.99)
    cyc = cfg.cycle_map()
    for phase in range(len(cyc)):
        g = np.random.default_rng(phase)

        class M:
            def root(self):
                return None, np.array([0.5, 0.5], np.float32), np.float32(0)

            def afterstate(self, sim, h, a):
                return None, np.array([0.5, 0.5], np.float32), np.float32(g.normal())

            def dynamics(self, sim, h, a):
                return None, np.array([0.5, 0.5], np.float32), np.float32(g.normal()), np.float32(g.normal())
        tree = O.search(cfg, M(), O.MTUniforms(phase), train=False, root_to_play=phase)
        for n in range(len(tree.visit)):
            d = tree.depth[n]
            assert tree.to_play[n] == to_play[phase, d]
            assert tree.is_chance[n] == bool((d >> 1) & 1)
            same = cyc[tree.to_play[0] % len(cyc)] == cyc[tree.to_play[n] % len(cyc)]
            assert sign[phase, d] == (1 if same else -1)


def test_dropin_surface_and_asserts():
    from stochastic_muzero_b200 import MinMaxStats, Monte_carlo_tree_search, Node, Player_cycle
    m = Monte_carlo_tree_search(pb_c_base=19652, pb_c_init=1.25, discount=0.997, root_dirichlet_alpha=0.25,
                                root_exploration_fraction=0.25, num_simulations=50, maxium_action_sample=2,
                                number_of_player=1, custom_loop=None)
    for k in ("pb_c_base", "pb_c_init", "discount", "root_dirichlet_alpha", "root_exploration_fraction",
              "num_simulations", "maxium_action_sample", "number_of_player", "custom_loop"):
        assert hasattr(m, k)
    assert m.cycle.global_step() == 0
    m.cycle.global_reset()
    with pytest.raises(AssertionError):
        Monte_carlo_tree_search(pb_c_base=1.5)
    with pytest.raises(AssertionError):
        Monte_carlo_tree_search(pb_c_init=1)
    with pytest.raises(AssertionError):
        Monte_carlo_tree_search(num_simulations=-1)
    n = Node(0.5)
    assert not n.expanded() and n.value() == 0 and n.to_play == -1 and n.is_chance is False
    mm = MinMaxStats()
    assert mm.normalize(3.0) == 3.0
    mm.update(1.0); mm.update(3.0)
    assert mm.normalize(2.0) == 0.5
    pc = Player_cycle(number_of_player=3)
    assert [pc.global_step() for _ in range(4)] == [0, 1, 2, 0]
    assert pc.proximate_player_step(2) == 0
    pc2 = Player_cycle(custom_loop="1>2>1>3")
    assert float(pc2.player_in_play(2)) == 1.0
    cfg = {"monte_carlo_tree_search": m.search_config()}
    m2 = Monte_carlo_tree_search.from_config(cfg)
    assert m2.search_config() == m.search_config()


def test_philox_known_answer():
    """Philox4x32-10 known-answer vectors from the Random123 distribution (kat_vectors)."""
    assert O.philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert O.philox4x32_10((0xffffffff,) * 4, (0xffffffff, 0xffffffff)) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert O.philox4x32_10((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


# ---------------------------------------------------------------------------------------------------
# the oracle's restatements of numpy's legacy RandomState algorithms, checked against numpy itself
# ---------------------------------------------------------------------------------------------------
def test_choice_restatement_matches_numpy_randomstate():
    """np.random.choice(n, size, p, replace=False) and choice(a, p=p): same picks AND same number of
    underlying random_sample() draws as numpy's own implementation, for random p / n / size."""
    g = np.random.default_rng(7)
    for trial in range(300):
        n = int(g.integers(1, 40))
        size = int(g.integers(1, n + 1))
        p = g.random(n).astype(np.float32) ** float(g.uniform(0.5, 6))
        p = O.normalise_policy(p)
        seed = int(g.integers(0, 2 ** 31))
        ref = np.random.RandomState(seed)
        expect = ref.choice(n, size, p=p, replace=False)
        rng = O.MTUniforms(seed)
        got = O.choice_without_replacement(n, size, p, rng)
        assert list(expect) == got, f"trial {trial}: n={n} size={size}"
        assert rng.rs.random_sample() == ref.random_sample(), "stream positions diverged (different draw count)"
        # with replacement, size=None (select_child's chance branch)
        ref2, rng2 = np.random.RandomState(seed), O.MTUniforms(seed)
        q = O.smoothed_chance_probs(p)
        e2 = ref2.choice(np.arange(n), p=q)
        g2 = int(np.searchsorted(O.choice_cdf(q), rng2.next(), side="right"))
        assert e2 == g2


def test_uniform_and_float32_sum_restatements():
    g = np.random.default_rng(8)
    for _ in range(200):
        seed = int(g.integers(0, 2 ** 31))
        a, b = np.random.RandomState(seed), np.random.RandomState(seed)
        assert a.uniform(low=1e-7, high=2e-7, size=1)[0] == np.float64(1e-7) + np.float64(2e-7 - 1e-7) * b.random_sample()
    # float32 add.reduce order (what the CUDA kernels and the oracle rely on): plain loop below 8, 8-way pairwise above
    def pairwise(x):
        n = len(x)
        if n < 8:
            r = np.float32(0)
            for v in x:
                r = np.float32(r + v)
            return r
        r = [np.float32(v) for v in x[:8]]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] = np.float32(r[j] + x[i + j])
            i += 8
        res = np.float32(np.float32(np.float32(r[0] + r[1]) + np.float32(r[2] + r[3])) +
                         np.float32(np.float32(r[4] + r[5]) + np.float32(r[6] + r[7])))
        while i < n:
            res = np.float32(res + x[i])
            i += 1
        return res
    for n in range(1, 33):
        for _ in range(40):
            x = g.random(n).astype(np.float32)
            assert x.sum() == pairwise(x), f"numpy float32 sum order changed for n={n}"
            assert x.mean() == np.float32(pairwise(x) / np.float32(n))


def test_markstein_division_is_exact_for_tabulated_divisors():
    """smz_div_r64 / smz_div_r32 (csrc/smz_common.cuh): q = RN(x*r); RN(q + (x - q*d)*r) with r = RN(1/d) must equal
    IEEE division for the small-integer divisors the tree kernels tabulate (visit counts, child counts) and for the
    min-max range divisor.  Checked here with exact rational arithmetic (an FMA rounds once)."""
    from fractions import Fraction
    import random

    def rn32(fr):
        f = np.float32(float(fr))
        cands = [f, np.nextafter(f, np.float32(np.inf)), np.nextafter(f, np.float32(-np.inf))]
        return np.float32(min(cands, key=lambda c: (abs(Fraction(float(c)) - fr), int(np.float32(c).view(np.uint32)) & 1)))

    rnd = random.Random(7)
    for _ in range(4000):
        d = rnd.randint(1, 64)
        x = rnd.uniform(0, 4) * 10 ** rnd.uniform(-6, 2)
        r = 1.0 / d
        q = x * r
        rem = float(Fraction(x) - Fraction(q) * d)                     # exact FMA, one rounding
        assert float(Fraction(q) + Fraction(rem) * Fraction(r)) == x / d
    for _ in range(4000):
        d = np.float32(rnd.randint(1, 64)) if rnd.random() < 0.5 else np.float32(rnd.uniform(1e-3, 60))
        x = np.float32(rnd.uniform(-50, 50))
        r = np.float32(1.0) / d
        q = np.float32(x * r)
        rem = rn32(Fraction(float(x)) - Fraction(float(q)) * Fraction(float(d)))
        got = rn32(Fraction(float(q)) + Fraction(float(rem)) * Fraction(float(r)))
        assert got == np.float32(x / d), (x, d)


def test_search_object_can_be_copied_and_pickled():
    """The reference hands the search object to ray workers (self_play.py:240-256): copies carry the nine
    parameters and the seed bookkeeping, never engine handles."""
    import copy
    import pickle
    from stochastic_muzero_b200 import Monte_carlo_tree_search
    m = Monte_carlo_tree_search(discount=0.997, num_simulations=7, maxium_action_sample=3, seed=11, tree_id_offset=64)
    m._engines["sentinel"] = object()          # stands for a live engine (ctypes handle: not picklable)
    for clone in (copy.deepcopy(m), pickle.loads(pickle.dumps(m))):
        assert clone._engines == {} and clone._weights_seen == {}
        assert (clone.discount, clone.num_simulations, clone.maxium_action_sample) == (0.997, 7, 3)
        assert clone._offset(16) == 64 and clone._base_seed() == 11
        assert clone.cycle.global_step() == 0
    m._engines.clear()
    assert Monte_carlo_tree_search(num_simulations=3, max_batch=32)._offset(8) == 0     # no process group: rank 0


def test_support_transform_fp32_is_quantised_near_zero():
    """Why value / reward scalars are compared with a looser bar than 1e-5: the reference evaluates
    inverse_transform_with_support (muzero_model.py:575-591) in float32, where sqrt(1 + 0.004 (|y| + 1.001)) - 1 cancels
    ~2.5 decimal digits.  Against float64 the reference's OWN formula is off by up to ~1e-4 near zero, and its output
    moves in whole quanta of ~1.2e-4."""
    f32 = np.float32

    def transform32(y):
        y = f32(y)
        inner = np.sqrt(f32(1) + f32(4) * f32(0.001) * (np.abs(y) + f32(1) + f32(0.001)), dtype=np.float32)
        return np.sign(y) * (((inner - f32(1)) / (f32(2) * f32(0.001))) ** 2 - f32(1))

    def transform64(y):
        y = np.float64(y)
        return np.sign(y) * (((np.sqrt(1 + 4 * 0.001 * (abs(y) + 1 + 0.001)) - 1) / (2 * 0.001)) ** 2 - 1)
    ys = np.random.default_rng(0).uniform(0.01, 0.5, 4000).astype(np.float32)
    err = np.array([abs(float(transform32(y)) - transform64(y)) for y in ys])
    assert 5e-5 < err.max() < 2.5e-4                       # the fp32 formula itself is ~1e-4 away from the exact value
    # as a function of y the fp32 output is a staircase: over a y interval of 1e-3 it takes a few dozen values only,
    # a whole quantum (~1.2e-4) apart — two evaluations whose y differ in the last bits agree exactly or by one quantum
    sweep = np.linspace(0.1, 0.101, 4001).astype(np.float32)
    vals = np.unique(np.array([float(transform32(y)) for y in sweep]))
    jumps = np.diff(vals)
    assert len(vals) < 40 and 5e-5 < jumps.min() and jumps.max() < 2.5e-4
