"""Pin the oracle against the live reference modules (only where /root/reference exists)."""
import os
import sys
import types

import numpy as np
import pytest
import torch

from oracle import cmgan_oracle as O

REF = "/root/reference/src"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, REF)
    if "pesq" not in sys.modules:
        stub = types.ModuleType("pesq")
        stub.pesq = lambda *a, **k: 0.0
        sys.modules["pesq"] = stub
    from models.generator import TSCNet
    from models.discriminator import Discriminator
    import utils as ref_utils
    sd = torch.load(os.path.join(REF, "best_ckpt", "ckpt"), map_location="cpu")
    m = TSCNet(64, 201)
    m.load_state_dict(sd)
    m.eval()
    return m, sd, Discriminator, ref_utils


def test_tscnet_random_input(ref):
    m, sd, _, _ = ref
    torch.manual_seed(3)
    x = torch.randn(1, 2, 23, 201).permute(0, 1, 2, 3) * 0.7
    with torch.no_grad():
        a = m(x)
        b = O.tscnet_forward(x, sd)
    for u, v in zip(a, b):
        assert (u - v).abs().max().item() < 3e-5


def test_stft_matches_torch():
    torch.manual_seed(0)
    x = torch.randn(3, 1700) * 0.1
    a = torch.view_as_real(torch.stft(x, 400, 100, window=torch.hamming_window(400), onesided=True, return_complex=True))
    assert (a - O.stft(x)).abs().max().item() < 2e-5
    y = torch.istft(torch.view_as_complex(a.contiguous()), 400, 100, window=torch.hamming_window(400), onesided=True)
    assert (y - O.istft(a)).abs().max().item() < 2e-6


def test_compress_matches(ref):
    _, _, _, U = ref
    torch.manual_seed(1)
    x = torch.randn(2, 201, 9, 2)
    x[0, 0, 0] = 0.0
    assert (U.power_compress(x) - O.power_compress(x)).abs().max().item() < 1e-6
    c = U.power_compress(x)
    assert (U.power_uncompress(c[:, 0:1], c[:, 1:2]) - O.power_uncompress(c[:, 0:1], c[:, 1:2])).abs().max().item() < 1e-5


def test_discriminator_train_mode(ref):
    _, _, Discriminator, _ = ref
    torch.manual_seed(11)
    D = Discriminator(ndf=16)
    dsd = {k: v.clone() for k, v in D.state_dict().items()}
    D.train()
    D.layers[15].p = 0.0
    x, y = torch.rand(3, 1, 201, 33), torch.rand(3, 1, 201, 33)
    a = D(x, y)
    b = O.discriminator_forward(x, y, dsd, training=True)
    assert (a - b).abs().max().item() < 1e-6
