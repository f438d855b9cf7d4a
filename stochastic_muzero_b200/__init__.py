"""Importable name of the package whose sources live in ``stochastic-muzero_b200/`` (a directory name
Python cannot import directly).  ``import stochastic_muzero_b200`` resolves sub-modules from there."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                                 "stochastic-muzero_b200"))

from .engine import SearchEngine, SmzError  # noqa: E402,F401
from .monte_carlo_tree_search import (  # noqa: E402,F401
    BatchedRoots, MinMaxStats, Monte_carlo_tree_search, Node, Player_cycle)
from .weights import (ModelShape, PackedModel, VisionShape, blob_layout, pack_vision_weights, pack_weights,
                      random_blob, vision_blob_layout)  # noqa: E402,F401
