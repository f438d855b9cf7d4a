"""Weight hand-off: the six MLP modules of a reference ``Muzero`` -> one flat fp32 blob.

This is the only coupling between the trainer and the search engine and the payload of the NCCL
weight broadcast.  Layout (torch ``Linear``: ``W[out, in]`` row-major, then
``b[out]``), nets in this order, hidden layers packed ONCE because the reference ties them
(neural_network_mlp_model.py:31-37 — the python list of modules is multiplied):

    repr : in(obs->H)  [mid(H->H) if L>0]  out(H->S)                       mlp:5-42
    pred : in(S->H)    [mid]               policy(H->A)  value(H->S)       mlp:47-83
    adyn : in(S+OH->H) [mid]               state(H->S)                     mlp:85-124 (reward head is dead code)
    apred: in(S->H)    [mid]               policy(H->C)  value(H->S)       mlp:127-163
    dyn  : in(S+OH->H) [mid]               reward(H->S)  state(H->S)       mlp:167-206
    enc  : in(obs->H)  [mid]               code(H->C)                      mlp:209-250

OH = max(A, C) is the one-hot width of both actions and chance codes (== A in the reference, where
the codebook size equals ``action_dimension``: muzero_model.py:508-509).
"""
from __future__ import annotations

import dataclasses
from typing import Dict, Tuple

import numpy as np


@dataclasses.dataclass(frozen=True)
class ModelShape:
    obs_dim: int
    action_dim: int
    chance_dim: int
    state_dim: int
    hidden_dim: int
    num_hidden_layers: int

    @property
    def onehot_dim(self) -> int:
        return max(self.action_dim, self.chance_dim)


def blob_layout(shape: ModelShape) -> Tuple[Dict[str, Tuple[int, Tuple[int, ...]]], int]:
    """name -> (offset in floats, tensor shape), total floats."""
    obs, A, C, S, H, L = (shape.obs_dim, shape.action_dim, shape.chance_dim, shape.state_dim,
                          shape.hidden_dim, shape.num_hidden_layers)
    OH = shape.onehot_dim
    entries = []

    def net(prefix, in_dim, heads):
        entries.extend([(f"{prefix}.in.w", (H, in_dim)), (f"{prefix}.in.b", (H,))])
        if L > 0:
            entries.extend([(f"{prefix}.mid.w", (H, H)), (f"{prefix}.mid.b", (H,))])
        for name, width in heads:
            entries.extend([(f"{prefix}.{name}.w", (width, H)), (f"{prefix}.{name}.b", (width,))])

    net("repr", obs, [("out", S)])
    net("pred", S, [("policy", A), ("value", S)])
    net("adyn", S + OH, [("state", S)])
    net("apred", S, [("policy", C), ("value", S)])
    net("dyn", S + OH, [("reward", S), ("state", S)])
    net("enc", obs, [("code", C)])
    layout, off = {}, 0
    for name, shp in entries:
        layout[name] = (off, shp)
        off += int(np.prod(shp))
    return layout, off


def random_blob(shape: ModelShape, seed: int = 0, std: float = 1.0 / 137.035999) -> np.ndarray:
    """Random-init weights as the reference's ``weights_init`` draws them: every Linear weight AND
    bias ~ N(0, 1/137.035999) (neural_network_mlp_model.py:495-508)."""
    _, total = blob_layout(shape)
    return (np.random.default_rng(seed).standard_normal(total) * std).astype(np.float32)


class PackedModel:
    """Weights already in hand-off form (blob + shape): accepted by ``run_batch`` / ``run`` wherever a
    reference ``Muzero`` is.  ``bump()`` after replacing ``blob`` in place to make engines re-upload."""
    model_structure = "mlp_model"

    def __init__(self, blob, shape: ModelShape):
        _, total = blob_layout(shape)
        blob = np.ascontiguousarray(blob, dtype=np.float32)
        if blob.size != total:
            raise ValueError(f"blob has {blob.size} floats, shape needs {total}")
        self.blob, self.shape, self.version = blob, shape, 0

    def bump(self):
        self.version += 1


def _unwrap(module):
    """muzero_model.py:360-367 wraps the modules in DataParallel when several GPUs are visible."""
    return module.module if module.__class__.__name__ == "DataParallel" else module


def _linears(sequential):
    import torch
    return [m for m in sequential if isinstance(m, torch.nn.Linear)]


def shape_of(model) -> ModelShape:
    """ModelShape of a reference-style ``Muzero`` (attributes set in muzero_model.py __init__)."""
    if isinstance(model, PackedModel):
        return model.shape
    rep = _linears(_unwrap(model.representation_function).state_norm)
    pol = _linears(_unwrap(model.prediction_function).policy)
    apol = _linears(_unwrap(model.afterstate_prediction_function).policy)
    H, obs = rep[0].weight.shape
    S = rep[-1].weight.shape[0]
    return ModelShape(obs_dim=int(obs), action_dim=int(pol[-1].weight.shape[0]),
                      chance_dim=int(apol[-1].weight.shape[0]), state_dim=int(S), hidden_dim=int(H),
                      num_hidden_layers=len(rep) - 2)


def weights_version(model) -> tuple:
    """Cheap change detector: torch bumps ``_version`` on every in-place update (optimizer steps)."""
    if isinstance(model, PackedModel):
        return (id(model), model.version)
    vers = []
    for name in ("representation", "prediction", "afterstate_dynamics", "afterstate_prediction", "dynamics",
                 "encoder"):
        mod = getattr(model, f"{name}_function")
        vers.append(id(mod))
        vers.extend((p.data_ptr(), p._version) for p in mod.parameters())
    return tuple(vers)


def pack_weights(model) -> Tuple[np.ndarray, ModelShape]:
    """Walk the six modules of a reference ``Muzero`` (or any object exposing the same six
    ``*_function`` Sequential stacks) and return (blob, shape)."""
    if isinstance(model, PackedModel):
        return model.blob, model.shape
    shape = shape_of(model)
    L = shape.num_hidden_layers
    parts = []

    def trunk(sequential):
        lin = _linears(sequential)
        if len(lin) != L + 2:
            raise ValueError(f"expected {L + 2} Linear layers in the stack, found {len(lin)}")
        if any(m is not lin[1] for m in lin[1:1 + L]):
            raise ValueError("hidden layers are not weight-tied; this layout packs one shared linear_mid")
        out = [lin[0].weight, lin[0].bias]
        if L > 0:
            out += [lin[1].weight, lin[1].bias]
        return out, lin[-1]

    rep = _unwrap(model.representation_function)
    t, out = trunk(rep.state_norm)
    parts += t + [out.weight, out.bias]
    for attr in ("prediction_function",):
        m = _unwrap(getattr(model, attr))
        t, pol = trunk(m.policy)
        _, val = trunk(m.value)
        parts += t + [pol.weight, pol.bias, val.weight, val.bias]
    m = _unwrap(model.afterstate_dynamics_function)
    t, st = trunk(m.next_state_normalized)
    parts += t + [st.weight, st.bias]
    m = _unwrap(model.afterstate_prediction_function)
    t, pol = trunk(m.policy)
    _, val = trunk(m.value)
    parts += t + [pol.weight, pol.bias, val.weight, val.bias]
    m = _unwrap(model.dynamics_function)
    t, rew = trunk(m.reward)
    _, st = trunk(m.next_state_normalized)
    parts += t + [rew.weight, rew.bias, st.weight, st.bias]
    m = _unwrap(model.encoder_function)
    t, code = trunk(m.encoder)
    parts += t + [code.weight, code.bias]
    blob = np.concatenate([p.detach().float().cpu().numpy().ravel() for p in parts]).astype(np.float32)
    _, total = blob_layout(shape)
    if blob.size != total:
        raise ValueError(f"packed {blob.size} floats, layout expects {total} (one-hot width mismatch?)")
    return blob, shape
