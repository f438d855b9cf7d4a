"""Helpers shared by the tests: load a committed golden tape batch (tests/golden/tree_*.npz)."""
import glob
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DUMP_KEYS = ("depth", "key", "visit", "value_sum", "reward", "prior", "is_chance", "to_play", "expanded")


def tree_cases():
    return sorted(os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "tree_*.npz")))


def net_cases():
    return sorted(os.path.basename(p)[4:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "net_*.npz")))


def load_tree_case(name):
    z = dict(np.load(os.path.join(GOLDEN_DIR, f"tree_{name}.npz")))
    z["config"] = json.loads(str(z.pop("config_json")))
    z["train"] = bool(z["train"])
    return z


def load_net_case(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, f"net_{name}.npz")))


def expected_dump(z, b):
    m = int(z["n_nodes"][b])
    d = {k: z["exp_" + k][b, :m] for k in DUMP_KEYS}
    d["minmax"] = z["exp_minmax"][b]
    return d


def assert_dump_equal(got, exp, what=""):
    """Bit-exact comparison of two canonical tree dumps (float columns compared as raw bits)."""
    for k in DUMP_KEYS:
        g, e = np.asarray(got[k]), np.asarray(exp[k])
        assert g.shape == e.shape, f"{what}: column {k} has {g.shape} nodes, expected {e.shape}"
        if e.dtype.kind == "f":
            g = g.astype(e.dtype)
            same = g.view(np.uint8).reshape(len(g), -1) == e.view(np.uint8).reshape(len(e), -1)
            ok = same.all(axis=1) | ((g == 0) & (e == 0))
            assert ok.all(), f"{what}: column {k} differs at node {np.flatnonzero(~ok)[:5]}: " \
                             f"{g[~ok][:5]!r} vs {e[~ok][:5]!r}"
        else:
            assert np.array_equal(g.astype(np.int64), e.astype(np.int64)), \
                f"{what}: column {k} differs at {np.flatnonzero(g != e)[:5]}"
    gm, em = np.asarray(got["minmax"], np.float32), np.asarray(exp["minmax"], np.float32)
    assert np.array_equal(gm, em), f"{what}: minmax {gm} vs {em}"


def vision_cases():
    return sorted(os.path.basename(p)[7:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "vision_*.npz")))


def load_vision_case(name):
    z = dict(np.load(os.path.join(GOLDEN_DIR, f"vision_{name}.npz")))
    z["obs"] = z["obs"].astype(np.float32)
    return z
