"""Module-level C entry points (cmgan_tscnet_*): the parameter table against the reference's state_dict contract on the CPU, the forward
against the nn.Module path and the float64 oracle on the GPU."""
import pytest
import torch

import cmgan_b200
from cmgan_b200 import module_abi


def test_param_table_matches_state_dict():
    sd = cmgan_b200.TSCNet(64, 201).state_dict()
    floats = [(k, v.numel()) for k, v in sd.items() if v.dtype.is_floating_point]
    table = module_abi.param_table()
    assert [(k, n) for k, _, n in table] == floats              # same keys, same order, same sizes (int64 num_batches_tracked excluded)
    end = 0
    for _, off, n in table:
        assert off % 4 == 0 and off >= end                      # 16-byte aligned, non-overlapping, ascending
        end = off + n
    from cmgan_b200._lib import lib
    assert lib().cdll.cmgan_tscnet_param_floats() >= end
    # the query needs no GPU and rejects shapes the network does not take
    assert module_abi.workspace_bytes(1, 81, 201, 1) > 0
    assert module_abi.workspace_bytes(4, 321, 201, 1) > module_abi.workspace_bytes(1, 321, 201, 1)
    with pytest.raises(RuntimeError):
        module_abi.workspace_bytes(1, 81, 200, 1)


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_tscnet_fwd_c_entry(precision):
    from cmgan_b200 import ops
    dev = torch.device("cuda", 0)
    ops.set_precision(precision)
    try:
        torch.manual_seed(3)
        model = cmgan_b200.TSCNet(64, 201).to(dev).eval()
        with torch.no_grad():
            for name, buf in model.named_buffers():             # non-trivial BatchNorm running statistics
                if name.endswith("running_mean"):
                    buf.normal_(0.0, 0.3)
                elif name.endswith("running_var"):
                    buf.uniform_(0.5, 1.5)
        x = torch.randn(2, 2, 201, 81, device=dev).permute(0, 1, 3, 2)          # the reference passes this permuted view (train.py:95)
        with torch.no_grad():
            ref_r, ref_i = model(x)
        flat = module_abi.pack_params(model.state_dict(), dev)
        p = 1 if precision == "tf32" else 0
        fr, fi = module_abi.tscnet_forward(flat, x, p)
        torch.cuda.synchronize()
        tol = 1e-6 * max(1.0, float(ref_r.abs().max()), float(ref_i.abs().max()))
        assert float((fr - ref_r).abs().max()) <= tol and float((fi - ref_i).abs().max()) <= tol       # same kernels, same order
        # status + message instead of an exception / crash when the workspace is too small
        small = torch.empty(1 << 20, dtype=torch.uint8, device=dev)
        with pytest.raises(RuntimeError, match="workspace too small"):
            module_abi.tscnet_forward(flat, x, p, workspace=small)
    finally:
        ops.set_precision("fp32")


@pytest.mark.gpu
def test_tscnet_fwd_c_entry_vs_oracle(g_weights):
    """shipped checkpoint, fp32: the C entry against the float64 oracle forward"""
    from oracle import cmgan_oracle as O
    dev = torch.device("cuda", 0)
    sd = {k: torch.as_tensor(v) for k, v in g_weights.items()}
    flat = module_abi.pack_params(sd, dev)
    torch.manual_seed(5)
    x = torch.randn(1, 2, 61, 201)
    fr, fi = module_abi.tscnet_forward(flat, x.to(dev), 0)
    P = {k: v.double() for k, v in sd.items()}
    rr, ri = O.tscnet_forward(x.double(), P)
    scale = max(float(rr.abs().max()), float(ri.abs().max()), 1.0)
    assert float((fr.cpu().double() - rr).abs().max()) <= 5e-5 * scale
    assert float((fi.cpu().double() - ri).abs().max()) <= 5e-5 * scale
