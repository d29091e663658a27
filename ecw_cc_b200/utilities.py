"""Host-side mirror of `utilities.subdiff` (reference utilities.py:26-73) over
the CUDA soft-threshold kernel (`ecw_subdiff`)."""
import numpy as np

from ._lib import lib, EcwError


def subdiff(eq, var, alpha, R_format=False):
    """Sub-gradient of the L1-regularised functional, element-wise:
    `var > 0 -> eq + alpha`, otherwise soft-threshold of `eq` (the reference's
    second loop runs over `var <= 0`, utilities.py:59-67)."""
    import torch
    if R_format:
        raise NotImplementedError("R_format conversion is broken in the reference as well (utilities.py:27)")
    is_t = isinstance(eq, torch.Tensor)
    shape = tuple(eq.shape)
    if shape != tuple(var.shape):  # utilities.py:39-40
        raise ValueError('equations and variables matrices must have the same shape')
    if not torch.cuda.is_available():
        raise EcwError("no CUDA device: subdiff has no CPU implementation in ecw_cc_b200")
    dev = eq.device if is_t else torch.device("cuda", torch.cuda.current_device())
    e = eq if is_t else torch.from_numpy(np.ascontiguousarray(eq, dtype=np.float64)).to(dev)
    v = var if isinstance(var, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(var, dtype=np.float64)).to(dev)
    e = e.contiguous()
    v = v.contiguous()
    out = torch.empty_like(e)
    st = torch.cuda.current_stream(dev).cuda_stream
    if lib.ecw_subdiff(e.data_ptr(), v.data_ptr(), float(alpha), out.data_ptr(), e.numel(), st) != 0:
        raise EcwError("ecw_subdiff failed")
    return out if is_t else out.cpu().numpy()
