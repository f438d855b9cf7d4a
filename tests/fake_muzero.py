"""A reference-SHAPED model for tests that must run where the reference checkout does not exist (the
GPU box): six torch modules with the attribute names and Sequential structure of
neural_network_mlp_model.py (tied hidden Linear, separate heads), filled from a weight blob, plus the
five ``*_function_inference`` methods of muzero_model.py:802-909 computed by the network oracle."""
import numpy as np
import torch
from torch import nn

from oracle import net_oracle as NO


def _stack(in_dim, H, L, out_dim, mid=None):
    first = nn.Linear(in_dim, H)
    mid = mid if mid is not None else nn.Linear(H, H)
    seq = [first, nn.ELU()] + [mid, nn.ELU()] * L + [nn.Linear(H, out_dim)]
    return nn.Sequential(*seq)


def _share_trunk(dst: nn.Sequential, src: nn.Sequential):
    for i in range(len(src) - 1):
        dst[i] = src[i]


class _Holder(nn.Module):
    pass


class FakeMuzero:
    model_structure = "mlp_model"

    def __init__(self, blob, obs, A, C, S, H, L):
        self.dims = (obs, A, C, S, H, L)
        self.net = NO.NetOracle(blob, obs, A, C, S, H, L)
        self.action_dimension, self.state_dimension = A, S
        OH = max(A, C)
        layout, _ = NO.blob_layout(obs, A, C, S, H, L)
        blob = np.asarray(blob, np.float32)

        def t(name):
            o, shp = layout[name]
            return torch.from_numpy(blob[o:o + int(np.prod(shp))].reshape(shp).copy())

        def fill(seq, prefix, head):
            lin = [m for m in seq if isinstance(m, nn.Linear)]
            with torch.no_grad():
                lin[0].weight.copy_(t(prefix + ".in.w")); lin[0].bias.copy_(t(prefix + ".in.b"))
                if L > 0:
                    lin[1].weight.copy_(t(prefix + ".mid.w")); lin[1].bias.copy_(t(prefix + ".mid.b"))
                lin[-1].weight.copy_(t(f"{prefix}.{head}.w")); lin[-1].bias.copy_(t(f"{prefix}.{head}.b"))

        def two_heads(prefix, in_dim, h1, w1, attr1, h2, w2, attr2):
            m = _Holder()
            a = _stack(in_dim, H, L, w1)
            b = _stack(in_dim, H, L, w2)
            _share_trunk(b, a)
            fill(a, prefix, h1); fill(b, prefix, h2)
            setattr(m, attr1, a); setattr(m, attr2, b)
            return m

        rep = _Holder(); rep.state_norm = _stack(obs, H, L, S); fill(rep.state_norm, "repr", "out")
        enc = _Holder(); enc.encoder = _stack(obs, H, L, C); fill(enc.encoder, "enc", "code")
        ady = _Holder(); ady.next_state_normalized = _stack(S + OH, H, L, S)
        fill(ady.next_state_normalized, "adyn", "state")
        self.representation_function = rep
        self.encoder_function = enc
        self.afterstate_dynamics_function = ady
        self.prediction_function = two_heads("pred", S, "policy", A, "policy", "value", S, "value")
        self.afterstate_prediction_function = two_heads("apred", S, "policy", C, "policy", "value", S, "value")
        self.dynamics_function = two_heads("dyn", S + OH, "reward", S, "reward", "state", S, "next_state_normalized")

    # the five inference methods (batch 1, numpy/torch on CPU like the reference)
    def representation_function_inference(self, obs):
        return torch.from_numpy(self.net.representation(np.asarray(obs, np.float32).reshape(1, -1)))

    def prediction_function_inference(self, h):
        p, v = self.net.prediction(np.asarray(h, np.float32))
        return p, v[0]

    def afterstate_prediction_function_inference(self, h):
        p, v = self.net.afterstate_prediction(np.asarray(h, np.float32))
        return p, v[0]

    def afterstate_dynamics_function_inference(self, h, a):
        return torch.from_numpy(self.net.afterstate_dynamics(np.asarray(h, np.float32), [int(a)]))

    def dynamics_function_inference(self, h, a):
        r, nh = self.net.dynamics(np.asarray(h, np.float32), [int(a)])
        return r[0], torch.from_numpy(nh)
