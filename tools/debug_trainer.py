import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch, torch.nn.functional as F
import cmgan_b200
from cmgan_b200 import training
from cmgan_b200.trainer import FusedTrainer
from oracle import cmgan_oracle as O
DEV = "cuda"
gw = O.load_weights_npz("tests/golden/weights_g.npz"); dw = O.load_weights_npz("tests/golden/weights_d.npz")
g = np.load("tests/golden/golden_small.npz")
clean = torch.from_numpy(g["grad_clean"]).to(DEV); noisy = torch.from_numpy(g["grad_noisy"]).to(DEV)
def models():
    m = cmgan_b200.TSCNet(64, 201); m.load_state_dict(gw, strict=True)
    d = cmgan_b200.Discriminator(16); d.load_state_dict(dw, strict=True)
    return m.to(DEV).eval(), d.to(DEV).eval()
keys = ["dense_encoder.conv_1.0.weight", "TSCB_2.time_conformer.ff1.fn.fn.net.0.weight", "mask_decoder.prelu_out.weight", "complex_decoder.conv.weight", "mask_decoder.final_conv.bias"]
for name, w in (("ri", (1, 0, 0, 0)), ("mag", (0, 1, 0, 0)), ("time", (0, 0, 1, 0)), ("gan", (0, 0, 0, 1)), ("all-but-gan", (0.1, 0.9, 0.2, 0))):
    m, d = models()
    go = training.forward_generator_step(m, clean, noisy)
    loss = training.generator_loss(go, clean, d if w[3] else None, weights=w)
    loss.backward()
    ref = {k: p.grad.clone() for k, p in m.named_parameters()}
    m2, d2 = models()
    t = FusedTrainer(m2, d2 if w[3] else None, weights=w)
    l2 = t.generator_step(clean, noisy, update=False)
    out = []
    for k in keys:
        a, b = dict(m2.named_parameters())[k].grad, ref[k]
        out.append(f"{(a - b).abs().max().item() / max(b.abs().max().item(), 1e-12):.2e}")
    print(name, f"loss {l2.item():.6f}/{loss.item():.6f}", out)
