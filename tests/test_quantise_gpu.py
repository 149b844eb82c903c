"""GPU quantiser vs the CAST macro of /root/reference/src/common.cpp:6 (bit-exact)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_quantise_device_bit_exact(oracle):
    import torch
    import denseflow_b200 as d
    rng = np.random.default_rng(0)
    h, w = 77, 203
    flow = (rng.standard_normal((h, w, 2)) * 15).astype(np.float32)
    # ties: v = L + (k+0.5)(H-L)/255 for every k, clamps, exact bounds
    for bound in (20, 32, 7):
        k = np.arange(255)
        ties = (-bound + (k + 0.5) * (2 * bound) / 255).astype(np.float32)
        flow[0, :255 if w >= 255 else w, 0] = ties[:min(w, 255)]
        flow[1, :8, 1] = [-bound, bound, -bound - 1e-3, bound + 1e-3, 0, -0.0, 1e30, -1e30]
        e = d.OpticalFlowDual_TVL1.create(0, 256, 256)
        qx, qy = e.quantise_device(torch.from_numpy(flow).cuda(), bound)
        ox, oy = oracle.quantise(flow, bound)
        assert np.array_equal(qx.cpu().numpy(), ox)
        assert np.array_equal(qy.cpu().numpy(), oy)
