"""Pins oracle/mcts_oracle.py against the committed outputs of the reference's own
Monte_carlo_tree_search.run (tests/golden/tree_*.npz, made by oracle/make_golden.py)."""
import numpy as np
import pytest

import golden_io
from oracle import mcts_oracle as O


def replay(z, b):
    c = z["config"]
    cfg = O.SearchConfig(**{k: c[k] for k in ("pb_c_base", "pb_c_init", "discount", "root_dirichlet_alpha",
                                               "root_exploration_fraction", "num_simulations",
                                               "maxium_action_sample", "number_of_player", "custom_loop")})
    rng = O.TapeUniforms(z["uniforms"][b, :z["n_uniforms"][b]])
    model = O.TapeModel(z["root_policy"][b], z["sim_policy"][b], z["sim_width"][b], z["sim_value"][b],
                        z["sim_reward"][b])
    tree = O.search(cfg, model, rng, train=z["train"], dirichlet=z["dirichlet"][b],
                    root_to_play=int(z["exp_root_to_play"][b]))
    return tree, rng


@pytest.mark.parametrize("name", golden_io.tree_cases())
def test_oracle_replays_reference_tape(name):
    z = golden_io.load_tree_case(name)
    for b in range(len(z["n_nodes"])):
        tree, rng = replay(z, b)
        assert rng.cursor == z["n_uniforms"][b], "oracle consumed a different number of uniform draws"
        golden_io.assert_dump_equal(tree.dump(), golden_io.expected_dump(z, b), f"{name}[{b}]")
        for s, keys in enumerate(tree.paths):
            exp = z["exp_paths"][b, s]
            assert list(exp[exp >= 0]) == keys, f"{name}[{b}] sim {s}: path differs"
