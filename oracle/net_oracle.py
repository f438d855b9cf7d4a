"""TEST INFRASTRUCTURE — CPU oracle for the per-simulation network step.  NOT PRODUCT CODE.

numpy float32 restatement of the reference's MLP inference path:
  * the six modules of /root/reference/neural_network_mlp_model.py (:5-42 Representation, :47-83
    Prediction, :85-124 Afterstate_dynamics, :127-163 Afterstate_prediction, :167-206 Dynamics,
    :209-250 Encoder) — Linear/ELU stacks whose L hidden layers share ONE weight (:31-37),
  * ``scale_to_bound_action`` (:349-357),
  * the inference facade of /root/reference/muzero_model.py: one-hot action (:496-509), softmax on
    the policy head (:837, :855), ``inverse_transform_with_support`` (:575-591).

Pinned by tests/test_oracle_golden.py against tests/golden/net_*.npz, which hold outputs of the
reference's own ``Muzero.*_function_inference`` on committed inputs (tolerance 2e-6: BLAS summation
order differs between torch and numpy; everything else is the same arithmetic).

Weight blob (the engine's hand-off format, documented in include/smz.h): one flat float32 array,
torch ``Linear`` layout ``W[out, in]`` row-major followed by ``b[out]``, in the order
  repr : in(obs->H) [mid(H->H) if L>0] out(H->S)
  pred : in(S->H)   [mid]              policy(H->A) value(H->S)
  adyn : in(S+OH->H)[mid]              state(H->S)
  apred: in(S->H)   [mid]              policy(H->C) value(H->S)
  dyn  : in(S+OH->H)[mid]              reward(H->S) state(H->S)
  enc  : in(obs->H) [mid]              code(H->C)
with OH = max(A, C) the one-hot width (== A in the reference, where C == A).
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def blob_layout(obs, A, C, S, H, L):
    """-> dict name -> (offset, shape), total size."""
    OH = max(A, C)
    spec = []

    def net(prefix, in_dim, heads):
        spec.append((f"{prefix}.in.w", (H, in_dim)))
        spec.append((f"{prefix}.in.b", (H,)))
        if L > 0:
            spec.append((f"{prefix}.mid.w", (H, H)))
            spec.append((f"{prefix}.mid.b", (H,)))
        for hname, width in heads:
            spec.append((f"{prefix}.{hname}.w", (width, H)))
            spec.append((f"{prefix}.{hname}.b", (width,)))

    net("repr", obs, [("out", S)])
    net("pred", S, [("policy", A), ("value", S)])
    net("adyn", S + OH, [("state", S)])
    net("apred", S, [("policy", C), ("value", S)])
    net("dyn", S + OH, [("reward", S), ("state", S)])
    net("enc", obs, [("code", C)])
    out, off = {}, 0
    for name, shape in spec:
        out[name] = (off, shape)
        off += int(np.prod(shape))
    return out, off


def elu(x):
    return np.where(x > 0, x, np.expm1(np.minimum(x, 0))).astype(np.float32)


def softmax(x):
    z = x - x.max(axis=-1, keepdims=True)
    e = np.exp(z).astype(np.float32)
    return (e / e.sum(axis=-1, keepdims=True)).astype(np.float32)


def scale_to_bound(x):
    """neural_network_mlp_model.py:349-357."""
    lo = x.min(axis=1, keepdims=True)
    hi = x.max(axis=1, keepdims=True)
    scale = (hi - lo).astype(np.float32)
    scale = np.where(scale < f32(1e-5), scale + f32(1e-5), scale).astype(np.float32)
    return ((x - lo) / scale).astype(np.float32)


def support_to_scalar(logits):
    """muzero_model.py:575-591 — softmax, expectation over the integer support, inverse of h(x)."""
    S = logits.shape[1]
    half = S // 2
    rem = int(2 * ((S / 2) - half))
    support = np.arange(-half, half + rem, dtype=np.float32)
    p = softmax(logits)
    y = (support[None, :] * p).sum(axis=1, dtype=np.float32)
    eps = f32(0.001)
    inner = np.sqrt(f32(1) + f32(4) * eps * (np.abs(y) + f32(1) + eps)).astype(np.float32)
    out = np.sign(y) * (((inner - f32(1)) / (f32(2) * eps)) ** 2 - f32(1))
    return out.astype(np.float32)


class NetOracle:
    def __init__(self, blob, obs, A, C, S, H, L):
        self.obs, self.A, self.C, self.S, self.H, self.L = obs, A, C, S, H, L
        self.OH = max(A, C)
        layout, total = blob_layout(obs, A, C, S, H, L)
        blob = np.asarray(blob, dtype=np.float32)
        assert blob.size == total, f"weight blob has {blob.size} floats, layout needs {total}"
        self.w = {k: blob[o:o + int(np.prod(s))].reshape(s) for k, (o, s) in layout.items()}

    def _lin(self, x, name):
        return (x @ self.w[name + ".w"].T + self.w[name + ".b"]).astype(np.float32)

    def _trunk(self, x, prefix):
        x = elu(self._lin(x, prefix + ".in"))
        for _ in range(self.L):
            x = elu(self._lin(x, prefix + ".mid"))
        return x

    def _onehot(self, idx):
        return np.eye(self.OH, dtype=np.float32)[np.asarray(idx, dtype=np.int64)]

    def representation(self, obs):
        return scale_to_bound(self._lin(self._trunk(obs.astype(np.float32), "repr"), "repr.out"))

    def _pred(self, h, prefix):
        t = self._trunk(h, prefix)
        return softmax(self._lin(t, prefix + ".policy")), support_to_scalar(self._lin(t, prefix + ".value"))

    def prediction(self, h):
        return self._pred(h, "pred")

    def afterstate_prediction(self, h):
        return self._pred(h, "apred")

    def afterstate_dynamics(self, h, action):
        t = self._trunk(np.concatenate([h, self._onehot(action)], axis=1), "adyn")
        return scale_to_bound(self._lin(t, "adyn.state"))

    def dynamics(self, h, code):
        t = self._trunk(np.concatenate([h, self._onehot(code)], axis=1), "dyn")
        return support_to_scalar(self._lin(t, "dyn.reward")), scale_to_bound(self._lin(t, "dyn.state"))

    def encoder(self, obs):
        probs = softmax(self._lin(self._trunk(obs.astype(np.float32), "enc"), "enc.code"))
        return probs, probs.argmax(axis=1).astype(np.int32)


class NetModel:
    """Adapter: drives oracle.mcts_oracle.search with the fp32 network oracle for one observation,
    logging what it returned (so the run can be replayed elsewhere as a tape)."""

    def __init__(self, net: NetOracle, obs_row):
        self.net, self.obs = net, np.asarray(obs_row, dtype=np.float32)[None, :]
        self.log = []

    def root(self):
        h = self.net.representation(self.obs)
        p, v = self.net.prediction(h)
        self.root_policy = p[0]
        return h, p[0], v[0]

    def afterstate(self, sim, parent_hidden, action):
        h = self.net.afterstate_dynamics(parent_hidden, [action])
        p, v = self.net.afterstate_prediction(h)
        self.log.append((0, p[0], v[0], f32(0)))
        return h, p[0], v[0]

    def dynamics(self, sim, parent_hidden, code):
        r, h = self.net.dynamics(parent_hidden, [code])
        p, v = self.net.prediction(h)
        self.log.append((1, p[0], v[0], r[0]))
        return h, p[0], v[0], r[0]
