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
    """Weights already in hand-off form (blob + shape; a ModelShape for the MLP family, a VisionShape for the
    vision family): accepted by ``run_batch`` / ``run`` wherever a
    reference ``Muzero`` is.  ``bump()`` after replacing ``blob`` in place to make engines re-upload."""
    model_structure = "mlp_model"

    def __init__(self, blob, shape: ModelShape):
        _, total = vision_blob_layout(shape) if isinstance(shape, VisionShape) else blob_layout(shape)
        blob = np.ascontiguousarray(blob, dtype=np.float32)
        if blob.size != total:
            raise ValueError(f"blob has {blob.size} floats, shape needs {total}")
        self.blob, self.shape, self.version = blob, shape, 0
        if isinstance(shape, VisionShape):
            self.model_structure = "vision_model"

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
        vers.extend((b.data_ptr(), b._version) for b in mod.buffers())      # BatchNorm running statistics
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


# ----------------------------------------------------------------------------------------------------
# vision (ResNet-v2) family — neural_network_vision_model.py, BASELINE config 5
# ----------------------------------------------------------------------------------------------------
VISION_HW = 7                       # hidden state [3, 7, 7]; the reference fixes the model input to 98x98x3
VISION_FLAT = 3 * VISION_HW * VISION_HW


@dataclasses.dataclass(frozen=True)
class VisionShape:
    action_dim: int
    state_dim: int          # categorical support size S of value / reward heads
    hidden_dim: int         # H of the MLP heads
    num_hidden_layers: int  # L: residual blocks per trunk AND tied hidden layers per MLP head


def vision_blob_layout(shape: VisionShape):
    """Flat fp32 hand-off layout of the vision family.  Residual block (vision:41-79): bn[4, c] = gamma, beta,
    running_mean, running_var; conv1[c,c,3,3] (used twice), conv3[c,c,3,3].  MLP head: in.w[H,147] in.b,
    [mid.w[H,H] mid.b if L>0 — tied], out.w[n,H] out.b.  Trunk residual blocks exist only when L > 0."""
    A, S, H, L = shape.action_dim, shape.state_dim, shape.hidden_dim, shape.num_hidden_layers
    spec = []

    def res(prefix, c):
        spec.extend([(f"{prefix}.bn", (4, c)), (f"{prefix}.conv1", (c, c, 3, 3)), (f"{prefix}.conv3", (c, c, 3, 3))])

    def mlp(prefix, n_out):
        spec.extend([(f"{prefix}.in.w", (H, VISION_FLAT)), (f"{prefix}.in.b", (H,))])
        if L > 0:
            spec.extend([(f"{prefix}.mid.w", (H, H)), (f"{prefix}.mid.b", (H,))])
        spec.extend([(f"{prefix}.out.w", (n_out, H)), (f"{prefix}.out.b", (n_out,))])

    spec.append(("repr.conv_in", (1, 3, 3, 3)))
    res("repr.res_in", 1)
    spec.append(("repr.conv_out", (3, 1, 3, 3)))
    res("repr.res_out", 3)
    res("repr.res_last", 3)
    for net in ("dyn", "adyn"):
        spec.extend([(f"{net}.conv", (3, 4, 3, 3)), (f"{net}.bn", (4, 3))])
        res(f"{net}.res", 3)
        if net == "dyn":
            spec.extend([("dyn.conv_reward.w", (3, 4)), ("dyn.conv_reward.b", (3,))])
            mlp("dyn.reward", S)
    for net in ("pred", "apred"):
        res(f"{net}.res", 3)
        spec.extend([(f"{net}.conv_value.w", (3, 3)), (f"{net}.conv_value.b", (3,))])
        mlp(f"{net}.value", S)
        spec.extend([(f"{net}.conv_policy.w", (3, 3)), (f"{net}.conv_policy.b", (3,))])
        mlp(f"{net}.policy", A)
    layout, off = {}, 0
    for name, shp in spec:
        layout[name] = (off, shp)
        off += int(np.prod(shp))
    return layout, off


def pack_vision_weights(model):
    """Walk the modules of a reference-style vision `Muzero` -> (blob, VisionShape).  With L == 0 the trunks
    have no residual block; the layout keeps the slots (identity: gamma=1, var=1, zero convs are NOT an
    identity for a v2 block, so L == 0 is rejected rather than faked)."""
    import torch

    def res_parts(rb):
        sc = rb.sequential_container
        bn, c1, c3 = sc[0], sc[2], sc[5]
        assert sc[8] is c1, "conv_1 is expected to be shared between positions 1 and 3 of the block"
        return [torch.stack([bn.weight, bn.bias, bn.running_mean, bn.running_var]), c1.weight, c3.weight]

    def bn_parts(bn):
        return [torch.stack([bn.weight, bn.bias, bn.running_mean, bn.running_var])]

    def mlp_parts(seq, L):
        lin = [m for m in seq if isinstance(m, torch.nn.Linear)]
        assert len(lin) == L + 2 and all(m is lin[1] for m in lin[1:1 + L])
        out = [lin[0].weight, lin[0].bias]
        if L > 0:
            out += [lin[1].weight, lin[1].bias]
        return out + [lin[-1].weight, lin[-1].bias]

    rep = _unwrap(model.representation_function)
    dyn, ady = _unwrap(model.dynamics_function), _unwrap(model.afterstate_dynamics_function)
    pre, apr = _unwrap(model.prediction_function), _unwrap(model.afterstate_prediction_function)
    mlp_v = pre.nn_value[2]
    lin = [m for m in mlp_v if isinstance(m, torch.nn.Linear)]
    L = len(lin) - 2
    if L < 1:
        raise ValueError("vision family with number_of_hidden_layer == 0 is not supported")
    shape = VisionShape(action_dim=int(pre.nn_policy[2][-1].weight.shape[0]), state_dim=int(lin[-1].weight.shape[0]),
                        hidden_dim=int(lin[0].weight.shape[0]), num_hidden_layers=L)
    ds = rep.sequential_downsampler[0].sequential_container
    parts = [ds[0].weight] + res_parts(ds[1]) + [ds[3].weight] + res_parts(ds[4]) + res_parts(rep.sequential_downsampler[1])
    for net in (dyn, ady):
        sc = net.sequential_container
        parts += [sc[0].weight] + bn_parts(sc[1]) + res_parts(sc[3])
        if net is dyn:
            parts += [net.sequential_reward[0].weight.reshape(3, 4), net.sequential_reward[0].bias]
            parts += mlp_parts(net.sequential_reward[2], L)
    for net in (pre, apr):
        parts += res_parts(net.resnet[0])
        parts += [net.nn_value[0].weight.reshape(3, 3), net.nn_value[0].bias] + mlp_parts(net.nn_value[2], L)
        parts += [net.nn_policy[0].weight.reshape(3, 3), net.nn_policy[0].bias] + mlp_parts(net.nn_policy[2], L)
    blob = np.concatenate([p.detach().float().cpu().numpy().ravel() for p in parts]).astype(np.float32)
    _, total = vision_blob_layout(shape)
    if blob.size != total:
        raise ValueError(f"packed {blob.size} floats, vision layout expects {total}")
    return blob, shape
