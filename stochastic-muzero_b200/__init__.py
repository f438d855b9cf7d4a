"""B200-native batched Stochastic-MuZero search: drop-in for the reference's monte_carlo_tree_search.py path.
Imported as ``stochastic_muzero_b200`` (see ../stochastic_muzero_b200.py)."""
from .engine import SearchEngine, SmzError  # noqa: F401
from .monte_carlo_tree_search import (  # noqa: F401
    BatchedRoots, MinMaxStats, Monte_carlo_tree_search, Node, Player_cycle, StaleSearchError)
from .weights import (ModelShape, PackedModel, VisionShape, blob_layout, pack_vision_weights, pack_weights,  # noqa: F401
                      random_blob, vision_blob_layout)
