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


def _series(which, approx, x, deriv, n):
    import ctypes as C
    torch = _lib.require_cuda()
    lib = _lib.load()
    is_tensor = isinstance(x, torch.Tensor)
    if is_tensor and x.is_cuda:
        dev, d_x = x.device, x.to(torch.float64).contiguous()
    else:
        dev = torch.device("cuda", torch.cuda.current_device())
        host = x.detach().cpu().numpy() if is_tensor else np.asarray(x, dtype=np.float64)
        d_x = torch.from_numpy(np.ascontiguousarray(host, dtype=np.float64)).to(dev)
    out = torch.empty_like(d_x)
    h = _tables.host_tables()
    coef = h.low_b if which == 0 else h.low_f
    with torch.cuda.device(dev):
        _lib.check(lib.ctb_jseries(which, approx, int(deriv), int(n), coef.ctypes.data_as(C.POINTER(C.c_double)),
                                   d_x.data_ptr(), out.data_ptr(), d_x.numel(),
                                   torch.cuda.current_stream(dev).cuda_stream), "ctb_jseries")
    if is_tensor:
        return out if x.is_cuda else out.cpu()
    res = out.cpu().numpy()
    return float(res) if res.ndim == 0 else res


def Jb_high(x, deriv=0, n=8):
    """Jb calculated using the high-x (low-T) expansion; x = m/T (finiteT.py:278-285)."""
    return _series(0, 0, x, deriv, n)


def Jf_high(x, deriv=0, n=8):
    """Jf calculated using the high-x (low-T) expansion; x = m/T (finiteT.py:288-296)."""
    return _series(1, 0, x, deriv, n)


def Jb_low(x, n=20):
    """Jb calculated using the low-x (high-T) expansion; x = m/T (finiteT.py:225-233)."""
    return _series(0, 1, x, 0, n)


def Jf_low(x, n=20):
    """Jf calculated using the low-x (high-T) expansion; x = m/T (finiteT.py:236-244)."""
    return _series(1, 1, x, 0, n)


def _exact(which, mode, x):
    """Jb / Jf from the defining integrals (finiteT.py:40-147) by device quadrature (ctb_jexact)."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    is_tensor = isinstance(x, torch.Tensor)
    if is_tensor and x.is_cuda:
        dev, d_x = x.device, x.to(torch.float64).contiguous()
    else:
        dev = torch.device("cuda", torch.cuda.current_device())
        host = x.detach().cpu().numpy() if is_tensor else np.asarray(x)
        if np.iscomplexobj(host):
            # the reference accepts x = i a for a negative mass squared (finiteT.py:45-52,75-81): theta = x^2 = -a^2
            if mode != _lib.EXACT_X or np.any((host.real != 0) & (host.imag != 0)):
                raise ValueError("complex x must be purely real or purely imaginary (and deriv = 0)")
            host, mode = (host * host).real, _lib.EXACT_THETA
        d_x = torch.from_numpy(np.ascontiguousarray(host, dtype=np.float64)).to(dev)
    out = torch.empty_like(d_x)
    with torch.cuda.device(dev):
        _lib.check(lib.ctb_jexact(which, mode, d_x.data_ptr(), out.data_ptr(), d_x.numel(),
                                  torch.cuda.current_stream(dev).cuda_stream), "ctb_jexact")
    if is_tensor:
        return out if x.is_cuda else out.cpu()
    res = out.cpu().numpy()
    return float(res) if res.ndim == 0 else res


def Jb_exact(x):
    """Jb calculated directly from the integral; x = m/T (finiteT.py:69-81,135-137)."""
    return _exact(0, _lib.EXACT_X, x)


def Jf_exact(x):
    """Jf calculated directly from the integral; x = m/T (finiteT.py:40-52,125-127)."""
    return _exact(1, _lib.EXACT_X, x)


def Jb_exact2(theta):
    """Jb calculated directly from the integral; input is theta = x^2, of either sign (finiteT.py:84-96)."""
    return _exact(0, _lib.EXACT_THETA, theta)


def Jf_exact2(theta):
    """Jf calculated directly from the integral; input is theta = x^2, of either sign (finiteT.py:54-66)."""
    return _exact(1, _lib.EXACT_THETA, theta)


def dJb_exact(x):
    """dJb/dx calculated directly from the integral (finiteT.py:109-111,145-147)."""
    return _exact(0, _lib.EXACT_DX, x)


def dJf_exact(x):
    """dJf/dx calculated directly from the integral (finiteT.py:104-106,140-142)."""
    return _exact(1, _lib.EXACT_DX, x)


def _front(which, x, approx, deriv, n):
    # same dispatch and the same ValueErrors as finiteT.Jb / Jf (finiteT.py:302-375)
    if approx == 'exact':
        if deriv == 0:
            return _exact(which, _lib.EXACT_X, x)
        elif deriv == 1:
            return _exact(which, _lib.EXACT_DX, x)
        raise ValueError("For approx=='exact', deriv must be 0 or 1.")
    elif approx == 'spline':
        return _eval("bf"[which], x, deriv)
    elif approx == 'low':
        if n > 100:
            raise ValueError("Must have n <= 100")
        if deriv == 0:
            if n > 50:
                raise ValueError("the coefficient table holds 50 terms (finiteT.py:207-222)")
            return _series(which, 1, x, 0, n)
        raise ValueError("For approx=='low', deriv must be 0.")
    elif approx == 'high':
        if deriv > 3:
            raise ValueError("For approx=='high', deriv must be 3 or less.")
        return _series(which, 0, x, deriv, n)
    raise ValueError("Unexpected value for 'approx'.")


def Jb(x, approx='high', deriv=0, n=8):
    """A shorthand for calling one of the Jb functions above; same signature and defaults as
    finiteT.Jb (finiteT.py:302-337).  For approx='spline' the argument is theta = (m/T)^2."""
    return _front(0, x, approx, deriv, n)


def Jf(x, approx='high', deriv=0, n=8):
    """Same for Jf (finiteT.py:340-375)."""
    return _front(1, x, approx, deriv, n)
