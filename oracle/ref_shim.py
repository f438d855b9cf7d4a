"""TEST INFRASTRUCTURE — container-only loader for the upstream reference.

Imports the read-only reference checkout (``$SMZ_REFERENCE`` or ``/root/reference``) in-process so
that the oracle restatement (``oracle/mcts_oracle.py``, ``oracle/net_oracle.py``) can be pinned
against the reference's own ``Monte_carlo_tree_search.run`` and so that ``oracle/make_golden.py``
can emit the golden tapes under ``tests/golden/``.

The reference imports ``gymnasium`` at muzero_model.py:11 but only touches
``gym.spaces.{Discrete, box.Box, tuple.Tuple}`` (muzero_model.py:484-494, :1008-1058); the package is
absent from this image, so a minimal stand-in is injected into ``sys.modules``.

Nothing here travels to the GPU box: ``/root/reference`` does not exist there.  ``available()`` is the
guard the tests use to skip the pin-against-reference cases.
"""
import os
import sys
import types
import warnings

import numpy as np

def _find_reference():
    """$SMZ_REFERENCE, then a driver-provided copy under baseline/_ref, then the container's mount."""
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for d in (os.environ.get("SMZ_REFERENCE"), os.path.join(repo, "baseline", "_ref"), "/root/reference"):
        if d and os.path.isfile(os.path.join(d, "monte_carlo_tree_search.py")):
            return d
    return os.environ.get("SMZ_REFERENCE", "/root/reference")


REFERENCE_DIR = _find_reference()


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "monte_carlo_tree_search.py"))


class Discrete:
    def __init__(self, n):
        self.n = n
        self.shape = ()
        self.dtype = np.int64


class Box:
    def __init__(self, low, high, shape, dtype=np.float32):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype


class Tuple(tuple):
    pass


def _inject_gymnasium_stub():
    if "gymnasium" in sys.modules:
        return
    gym = types.ModuleType("gymnasium")
    spaces = types.ModuleType("gymnasium.spaces")
    box = types.ModuleType("gymnasium.spaces.box")
    tup = types.ModuleType("gymnasium.spaces.tuple")
    box.Box, tup.Tuple = Box, Tuple
    spaces.Discrete, spaces.Box, spaces.Tuple, spaces.box, spaces.tuple = Discrete, Box, Tuple, box, tup
    gym.spaces = spaces
    sys.modules.update({"gymnasium": gym, "gymnasium.spaces": spaces,
                        "gymnasium.spaces.box": box, "gymnasium.spaces.tuple": tup})


_loaded = {}


def load():
    """Return (monte_carlo_tree_search module, muzero_model module) of the reference."""
    if _loaded:
        return _loaded["mcts"], _loaded["model"]
    if not available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_DIR}")
    _inject_gymnasium_stub()
    # The reference is a flat directory of modules; our own package also ships a drop-in module
    # named monte_carlo_tree_search, so import the reference's by explicit path precedence.
    sys.path.insert(0, REFERENCE_DIR)
    try:
        for name in ("monte_carlo_tree_search", "muzero_model", "neural_network_mlp_model"):
            if name in sys.modules and not getattr(sys.modules[name], "__file__", "").startswith(REFERENCE_DIR):
                del sys.modules[name]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            import monte_carlo_tree_search as ref_mcts
            import muzero_model as ref_model
    finally:
        sys.path.remove(REFERENCE_DIR)
        sys.path.append(REFERENCE_DIR)  # lazy `__import__` of network modules (muzero_model.py:308-318)
    _loaded["mcts"], _loaded["model"] = ref_mcts, ref_model
    return ref_mcts, ref_model


def make_muzero(obs_dim=4, action_dim=2, state_dim=61, hidden_dim=126, n_hidden=4, seed=0):
    """Reference ``Muzero`` MLP with ``weights_init`` random init (neural_network_mlp_model.py:495-508)."""
    import torch
    _, ref_model = load()
    torch.manual_seed(seed)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return ref_model.Muzero(model_structure="mlp_model",
                                observation_space_dimensions=Box(-1, 1, (obs_dim,)),
                                action_space_dimensions=Discrete(action_dim),
                                state_space_dimensions=state_dim,
                                hidden_layer_dimensions=hidden_dim,
                                number_of_hidden_layer=n_hidden,
                                device="cpu", use_amp=False)


def load_checkpoint(tag=450):
    """Load the six pickled modules of a shipped checkpoint (muzero_model.py:952-996) into a Muzero."""
    import json
    import torch
    load()
    d = os.path.join(REFERENCE_DIR, "model_checkpoint")
    with open(os.path.join(d, f"{tag}_muzero_init_variables.json")) as f:
        iv = json.load(f)
    mz = make_muzero(obs_dim=int(iv["observation_space_dimensions"]), action_dim=len(iv["action_map"]),
                     state_dim=iv["state_space_dimensions"], hidden_dim=iv["hidden_layer_dimensions"],
                     n_hidden=iv["number_of_hidden_layer"])
    for name in ("representation", "prediction", "afterstate_prediction", "afterstate_dynamics",
                 "dynamics", "encoder"):
        mod = torch.load(os.path.join(d, f"{tag}_muzero_{name}_function.pt"), weights_only=False,
                         map_location="cpu")
        setattr(mz, f"{name}_function", mod.float())
    return mz
