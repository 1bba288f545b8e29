"""GPU counterparts of cosmoTransitions.finiteT.Jb_spline / Jf_spline (finiteT.py:167-174,197-204)
and of the spline branch of the Jb / Jf front ends (finiteT.py:302-375).

Input: theta = (m/T)^2 as a numpy array (any shape) or a torch tensor; output has the same shape and
lives where the input lives (numpy in -> numpy out through the current CUDA device)."""
import numpy as np

from . import _lib, _tables


def _eval(which, X, n):
    torch = _lib.require_cuda()
    lib = _lib.load()
    if n not in (0, 1, 2, 3):
        raise ValueError("deriv must be 0..3 for approx='spline'")
    is_tensor = isinstance(X, torch.Tensor)
    if is_tensor and X.is_cuda:
        dev = X.device
        d_x = X.to(torch.float64).contiguous()
    else:
        dev = torch.device("cuda", torch.cuda.current_device())
        host = X.detach().cpu().numpy() if is_tensor else np.asarray(X, dtype=np.float64)
        d_x = torch.from_numpy(np.ascontiguousarray(host, dtype=np.float64)).to(dev)
    tabs = _tables.device_tables(dev)
    out = torch.empty_like(d_x)
    with torch.cuda.device(dev):
        _lib.check(lib.ctb_jspline(tabs.jb if which == "b" else tabs.jf, n, d_x.data_ptr(), out.data_ptr(),
                                   d_x.numel(), torch.cuda.current_stream(dev).cuda_stream), "ctb_jspline")
    if is_tensor:
        return out if X.is_cuda else out.cpu()
    res = out.cpu().numpy()
    return res


def Jb_spline(X, n=0):
    """Jb interpolated from the saved spline. Input is (m/T)^2."""
    return _eval("b", X, n)


def Jf_spline(X, n=0):
    """Jf interpolated from the saved spline. Input is (m/T)^2."""
    return _eval("f", X, n)


def Jb(x, approx='spline', deriv=0, n=8):
    """Front end mirroring finiteT.Jb (finiteT.py:302-337) for approx='spline' (the variant
    generic_potential uses, generic_potential.py:20-21)."""
    if approx == 'spline':
        return Jb_spline(x, deriv)
    raise NotImplementedError("approx=%r: only the spline branch runs on the device (SURVEY.md s.8f row 4)" % approx)


def Jf(x, approx='spline', deriv=0, n=8):
    if approx == 'spline':
        return Jf_spline(x, deriv)
    raise NotImplementedError("approx=%r: only the spline branch runs on the device (SURVEY.md s.8f row 4)" % approx)
