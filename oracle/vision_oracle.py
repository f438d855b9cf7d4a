"""TEST INFRASTRUCTURE — CPU oracle for the vision (ResNet-v2) network family.  NOT PRODUCT CODE.

numpy float32 restatement of the inference forward of /root/reference/neural_network_vision_model.py as
the search path uses it (BASELINE config 5): Residual_block v2 with ONE shared BatchNorm2d and a conv used
twice (:41-79), Down_sample (:81-119), Representation (:122-158), Dynamics / Afterstate_dynamics (:161-226,
:373-430), Prediction / Afterstate_prediction (:229-296, :433-492), channel-wise scale_to_bound_action
(:495-503), plus the facade of muzero_model.py for RGB models: the action enters as a constant plane
(a+1)/A (:511-522), softmax on the policy, inverse_transform_with_support on value / reward.
BatchNorm is in eval mode (running statistics), exactly as the `*_inference` methods run it.

Pinned by tests/test_oracle_golden.py against tests/golden/net_vision*.npz (outputs of the reference's
own inference methods on committed inputs; BN statistics randomised so the fold is exercised).

Vision weight blob (flat float32, order = blob_layout below).  A residual block is
  bn{gamma, beta, mean, var}[c]  conv1[c,c,3,3]  conv3[c,c,3,3]
and an MLP head is  in.w[H,147] in.b[H]  [mid.w[H,H] mid.b[H] if L>0]  out.w[n,H] out.b[n]  (hidden layer
tied: the python list of modules is multiplied in the reference).
"""
from __future__ import annotations

import numpy as np

from oracle.net_oracle import softmax, support_to_scalar

f32 = np.float32
BN_EPS = 1e-5
HW = 7                      # hidden state is [3, 7, 7]
FLAT = 3 * HW * HW          # 147


def blob_layout(A, S, H, L):
    spec = []

    def res(prefix, c):
        spec.extend([(f"{prefix}.bn", (4, c)), (f"{prefix}.conv1", (c, c, 3, 3)), (f"{prefix}.conv3", (c, c, 3, 3))])

    def mlp(prefix, n_out):
        spec.extend([(f"{prefix}.in.w", (H, FLAT)), (f"{prefix}.in.b", (H,))])
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
    out, off = {}, 0
    for name, shape in spec:
        out[name] = (off, shape)
        off += int(np.prod(shape))
    return out, off


def conv3x3(x, w, stride=1):
    """x [B,Ci,H,W], w [Co,Ci,3,3], padding 1, no bias."""
    B, Ci, Hh, Ww = x.shape
    Ho, Wo = (Hh + 2 - 3) // stride + 1, (Ww + 2 - 3) // stride + 1
    xp = np.zeros((B, Ci, Hh + 2, Ww + 2), np.float32)
    xp[:, :, 1:-1, 1:-1] = x
    out = np.zeros((B, w.shape[0], Ho, Wo), np.float32)
    for ky in range(3):
        for kx in range(3):
            patch = xp[:, :, ky:ky + stride * (Ho - 1) + 1:stride, kx:kx + stride * (Wo - 1) + 1:stride]
            out += np.einsum("bchw,oc->bohw", patch, w[:, :, ky, kx]).astype(np.float32)
    return out


def conv1x1(x, w, b):
    return (np.einsum("bchw,oc->bohw", x, w) + b[None, :, None, None]).astype(np.float32)


def avgpool3s2(x):
    """AvgPool2d(kernel 3, stride 2, padding 1), count_include_pad=True (torch default)."""
    B, C, Hh, Ww = x.shape
    Ho, Wo = (Hh + 2 - 3) // 2 + 1, (Ww + 2 - 3) // 2 + 1
    xp = np.zeros((B, C, Hh + 2, Ww + 2), np.float32)
    xp[:, :, 1:-1, 1:-1] = x
    out = np.zeros((B, C, Ho, Wo), np.float32)
    for ky in range(3):
        for kx in range(3):
            out += xp[:, :, ky:ky + 2 * (Ho - 1) + 1:2, kx:kx + 2 * (Wo - 1) + 1:2]
    return (out / f32(9)).astype(np.float32)


def bn_relu(x, bn, relu=True):
    gamma, beta, mean, var = bn
    y = (x - mean[None, :, None, None]) / np.sqrt(var[None, :, None, None] + f32(BN_EPS)) * \
        gamma[None, :, None, None] + beta[None, :, None, None]
    return np.maximum(y, 0).astype(np.float32) if relu else y.astype(np.float32)


def scale_to_bound_channels(x):
    """vision:495-503 — min/max over dim 1 (the 3 channels of every pixel)."""
    lo, hi = x.min(1, keepdims=True), x.max(1, keepdims=True)
    sc = (hi - lo).astype(np.float32)
    sc = np.where(sc < f32(1e-5), sc + f32(1e-5), sc)
    return ((x - lo) / sc).astype(np.float32)


class VisionOracle:
    def __init__(self, blob, A, S, H, L):
        self.A, self.S, self.H, self.L = A, S, H, L
        layout, total = blob_layout(A, S, H, L)
        blob = np.asarray(blob, np.float32)
        assert blob.size == total, f"vision blob has {blob.size} floats, layout needs {total}"
        self.w = {k: blob[o:o + int(np.prod(s))].reshape(s) for k, (o, s) in layout.items()}

    def _res(self, x, p):
        bn, c1, c3 = self.w[p + ".bn"], self.w[p + ".conv1"], self.w[p + ".conv3"]
        y = conv3x3(bn_relu(x, bn), c1)
        y = conv3x3(bn_relu(y, bn), c3)
        y = conv3x3(bn_relu(y, bn), c1)
        return (y + x).astype(np.float32)

    def _mlp(self, x, p):
        x = np.maximum(x @ self.w[p + ".in.w"].T + self.w[p + ".in.b"], 0).astype(np.float32)
        for _ in range(self.L):
            x = np.maximum(x @ self.w[p + ".mid.w"].T + self.w[p + ".mid.b"], 0).astype(np.float32)
        return (x @ self.w[p + ".out.w"].T + self.w[p + ".out.b"]).astype(np.float32)

    def representation(self, obs):
        """obs [B,3,98,98] -> hidden [B,3,7,7]"""
        x = conv3x3(obs.astype(np.float32), self.w["repr.conv_in"], stride=2)
        x = self._res(self._res(x, "repr.res_in"), "repr.res_in")
        x = conv3x3(x, self.w["repr.conv_out"], stride=2)
        x = self._res(self._res(x, "repr.res_out"), "repr.res_out")
        x = avgpool3s2(x)
        for _ in range(3):
            x = self._res(x, "repr.res_out")
        x = avgpool3s2(x)
        x = self._res(x, "repr.res_last")
        return scale_to_bound_channels(x)

    def _with_action(self, h, idx):
        plane = ((np.asarray(idx, np.float32) + f32(1)) / f32(self.A)).astype(np.float32)
        return np.concatenate([h, np.broadcast_to(plane[:, None, None, None], (h.shape[0], 1, HW, HW))], 1).astype(np.float32)

    def _dyn_trunk(self, x, net):
        y = bn_relu(conv3x3(x, self.w[net + ".conv"]), self.w[net + ".bn"])
        for _ in range(self.L):
            y = self._res(y, net + ".res")
        return scale_to_bound_channels(np.maximum(y, 0))

    def afterstate_dynamics(self, h, action):
        return self._dyn_trunk(self._with_action(h, action), "adyn")

    def dynamics(self, h, code):
        x = self._with_action(h, code)
        r = conv1x1(x, self.w["dyn.conv_reward.w"], self.w["dyn.conv_reward.b"]).reshape(x.shape[0], -1)
        return support_to_scalar(self._mlp(r, "dyn.reward")), self._dyn_trunk(x, "dyn")

    def _pred(self, h, net):
        y = h
        for _ in range(self.L):
            y = self._res(y, net + ".res")
        v = conv1x1(y, self.w[net + ".conv_value.w"], self.w[net + ".conv_value.b"]).reshape(h.shape[0], -1)
        p = conv1x1(y, self.w[net + ".conv_policy.w"], self.w[net + ".conv_policy.b"]).reshape(h.shape[0], -1)
        return softmax(self._mlp(p, net + ".policy")), support_to_scalar(self._mlp(v, net + ".value"))

    def prediction(self, h):
        return self._pred(h, "pred")

    def afterstate_prediction(self, h):
        return self._pred(h, "apred")
