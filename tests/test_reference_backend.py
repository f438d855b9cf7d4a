"""Container-only checks (skipped where the reference checkout does not exist, e.g. on the GPU box):
the torch re-statement of the inference facade in batched_model.ReferenceModuleBackend against the
reference's own batch-1 `*_function_inference` methods, for the MLP and the vision model families."""
import warnings

import numpy as np
import pytest
import torch

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference checkout not present")


def _vision_muzero(A=4, S=61, H=32, L=2, seed=0):
    _, ref_model = ref_shim.load()
    torch.manual_seed(seed)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return ref_model.Muzero(model_structure="vision_model",
                                observation_space_dimensions=ref_shim.Box(0, 1, (98, 98, 3)),
                                action_space_dimensions=ref_shim.Discrete(A), state_space_dimensions=S,
                                hidden_layer_dimensions=H, number_of_hidden_layer=L, device="cpu", use_amp=False)


@pytest.mark.parametrize("family", ["mlp", "vision"])
def test_reference_module_backend_matches_reference_inference(family):
    from stochastic_muzero_b200.batched_model import ReferenceModuleBackend
    torch.set_num_threads(1)
    if family == "mlp":
        mz = ref_shim.make_muzero(obs_dim=5, action_dim=3, state_dim=11, hidden_dim=14, n_hidden=2, seed=1)
        obs = torch.randn(6, 5)
    else:
        mz = _vision_muzero()
        obs = torch.rand(3, 3, 98, 98)
    # vision: scale_to_bound_action normalises over the 3 CHANNELS of each pixel (vision:495-503); when the
    # three values nearly coincide the division amplifies fp32 noise (batched vs batch-1 conv algorithms)
    hid_tol = 1e-5 if family == "mlp" else 2e-4
    be = ReferenceModuleBackend(mz, device="cpu")
    A = mz.action_dimension
    acts = torch.arange(obs.shape[0]) % A
    h = be.representation(obs)
    p, v = be.prediction(h)
    ah = be.afterstate_dynamics(h, acts)
    ap, av = be.afterstate_prediction(ah)
    r, dh = be.dynamics(ah, acts)
    with torch.no_grad():
        for i in range(obs.shape[0]):
            rh = mz.representation_function_inference(obs[i:i + 1])
            np.testing.assert_allclose(h[i:i + 1].numpy(), rh.numpy(), atol=hid_tol)
            rp, rv = mz.prediction_function_inference(rh)
            np.testing.assert_allclose(p[i].numpy(), rp[0], atol=1e-5)
            np.testing.assert_allclose(v[i].item(), rv, atol=1e-5, rtol=5e-5)
            rah = mz.afterstate_dynamics_function_inference(rh, int(acts[i]))
            np.testing.assert_allclose(ah[i:i + 1].numpy(), rah.numpy(), atol=hid_tol)
            rap, rav = mz.afterstate_prediction_function_inference(rah)
            np.testing.assert_allclose(ap[i].numpy(), rap[0], atol=1e-5)
            np.testing.assert_allclose(av[i].item(), rav, atol=1e-5, rtol=5e-5)
            rr, rdh = mz.dynamics_function_inference(rah, int(acts[i]))
            np.testing.assert_allclose(dh[i:i + 1].numpy(), rdh.numpy(), atol=hid_tol)
            np.testing.assert_allclose(r[i].item(), rr, atol=1e-5, rtol=5e-5)
