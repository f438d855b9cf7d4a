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


# ---------------------------------------------------------------------------------------------------
# network oracle vs the reference's own *_inference outputs
# ---------------------------------------------------------------------------------------------------
NET_TOL = 2e-6   # numpy vs torch BLAS summation order; same arithmetic otherwise


@pytest.mark.parametrize("name", golden_io.net_cases())
def test_net_oracle_matches_reference_inference(name):
    from oracle import net_oracle as NO
    z = golden_io.load_net_case(name)
    net = NO.NetOracle(z["weights"], *[int(v) for v in z["dims"]])
    h = net.representation(z["obs"])
    np.testing.assert_allclose(h, z["repr_h"], atol=NET_TOL, rtol=0)
    p, v = net.prediction(z["repr_h"])
    np.testing.assert_allclose(p, z["pred_policy"], atol=NET_TOL, rtol=0)
    np.testing.assert_allclose(v, z["pred_value"], atol=NET_TOL, rtol=1e-5)
    ah = net.afterstate_dynamics(z["repr_h"], z["actions"])
    np.testing.assert_allclose(ah, z["adyn_h"], atol=NET_TOL, rtol=0)
    ap, av = net.afterstate_prediction(z["adyn_h"])
    np.testing.assert_allclose(ap, z["apred_policy"], atol=NET_TOL, rtol=0)
    np.testing.assert_allclose(av, z["apred_value"], atol=NET_TOL, rtol=1e-5)
    r, dh = net.dynamics(z["adyn_h"], z["actions"])
    np.testing.assert_allclose(dh, z["dyn_h"], atol=NET_TOL, rtol=0)
    np.testing.assert_allclose(r, z["dyn_reward"], atol=NET_TOL, rtol=1e-5)
    dp, dv = net.prediction(z["dyn_h"])
    np.testing.assert_allclose(dp, z["dpred_policy"], atol=NET_TOL, rtol=0)
    np.testing.assert_allclose(dv, z["dpred_value"], atol=NET_TOL, rtol=1e-5)
    probs, code = net.encoder(z["obs"])
    np.testing.assert_allclose(probs, z["enc_probs"], atol=NET_TOL, rtol=0)
    margin = np.abs(z["enc_probs"][:, :1] - z["enc_probs"]).max(axis=1) if z["enc_probs"].shape[1] == 2 else None
    assert np.array_equal(code, z["enc_code"]) or margin is not None and (margin[code != z["enc_code"]] < 1e-5).all()


# ---------------------------------------------------------------------------------------------------
# read-out / action selection oracle vs the reference's Game methods (game.py:179-216)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", golden_io.tree_cases())
def test_readout_oracle_matches_reference_game_methods(name):
    from oracle import readout_oracle as RO
    z = golden_io.load_tree_case(name)
    A = z["config"]["action_dim"]
    for b in range(len(z["n_nodes"])):
        e = golden_io.expected_dump(z, b)
        kids = np.flatnonzero(e["depth"] == 1)
        visits, priors = e["visit"][kids], e["prior"][kids]
        assert np.array_equal(RO.stored_policy(visits, priors), z["readout_stored"][b])
        for t_i, T in enumerate(z["readout_temperature"]):
            pol = RO.step_policy(visits, priors, T)
            assert np.array_equal(pol, z["readout_policy"][b, t_i])
            idx, sampled = RO.select_action(pol, T, z["readout_u"][b, t_i])
            assert sampled == bool(z["readout_sampled"][b, t_i]) and idx == z["readout_index"][b, t_i]


# ---------------------------------------------------------------------------------------------------
# vision (ResNet-v2) network oracle vs the reference's inference outputs (BASELINE config 5 family)
# ---------------------------------------------------------------------------------------------------
# scale_to_bound_action normalises over the 3 channels of each pixel: when they nearly coincide the
# division amplifies fp32 summation-order noise, hence the looser bar on hidden states
VISION_HIDDEN_ATOL = 2e-4


@pytest.mark.parametrize("name", golden_io.vision_cases())
def test_vision_oracle_matches_reference_inference(name):
    from oracle import vision_oracle as VO
    z = golden_io.load_vision_case(name)
    A, S, H, L = [int(v) for v in z["dims"]]
    net = VO.VisionOracle(z["weights"], A, S, H, L)
    np.testing.assert_allclose(net.representation(z["obs"]), z["repr_h"], atol=VISION_HIDDEN_ATOL)
    p, v = net.prediction(z["repr_h"])
    np.testing.assert_allclose(p, z["pred_policy"], atol=5e-6)
    np.testing.assert_allclose(v, z["pred_value"], atol=1e-5, rtol=5e-5)
    np.testing.assert_allclose(net.afterstate_dynamics(z["repr_h"], z["actions"]), z["adyn_h"], atol=VISION_HIDDEN_ATOL)
    ap, av = net.afterstate_prediction(z["adyn_h"])
    np.testing.assert_allclose(ap, z["apred_policy"], atol=5e-6)
    np.testing.assert_allclose(av, z["apred_value"], atol=1e-5, rtol=5e-5)
    r, dh = net.dynamics(z["adyn_h"], z["actions"])
    np.testing.assert_allclose(dh, z["dyn_h"], atol=VISION_HIDDEN_ATOL)
    np.testing.assert_allclose(r, z["dyn_reward"], atol=1e-5, rtol=5e-5)
    dp, dv = net.prediction(z["dyn_h"])
    np.testing.assert_allclose(dp, z["dpred_policy"], atol=5e-6)
    np.testing.assert_allclose(dv, z["dpred_value"], atol=1e-5, rtol=5e-5)
