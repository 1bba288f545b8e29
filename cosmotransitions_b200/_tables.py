"""Jb / Jf spline tables: the reference's own 2x1000 exact-integral samples and scipy's splrep
(t, c, k=3) (finiteT.py:151-166,178-196), frozen in data/finiteT_tables.npz by oracle/gen_golden.py,
converted here (numpy, once) to one cubic per knot interval and uploaded per device."""
import math
import os

import numpy as np

from . import _lib

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "finiteT_tables.npz")


def bspline_to_pp(t, c, k=3):
    """B-spline (t, c, k) -> (breaks[nint+1], coef[nint, k+1]) with
    S(x) = sum_m coef[j, m] (x - breaks[j])^m on [breaks[j], breaks[j+1])."""
    t = np.asarray(t, dtype=np.float64)
    n = len(t)
    derivs = [np.asarray(c, dtype=np.float64)[:n - k - 1].copy()]
    for m in range(1, k + 1):
        p = k - m + 1
        prev = derivs[-1]
        tk = t[m - 1:n - (m - 1)]
        denom = tk[p + 1:p + len(prev)] - tk[1:len(prev)]
        d = np.zeros(len(prev) - 1)
        nz = denom > 0
        d[nz] = p * (prev[1:] - prev[:-1])[nz] / denom[nz]
        derivs.append(d)
    ls = [l for l in range(k, n - k - 1) if t[l] < t[l + 1]]
    brk = np.array([t[l] for l in ls] + [t[ls[-1] + 1]])
    coef = np.zeros((len(ls), k + 1))
    for m in range(k + 1):
        p = k - m
        tk = t[m:n - m]
        cm = derivs[m]
        for row, l in enumerate(ls):
            lm = l - m
            x = t[l]
            h = [1.0]
            for j in range(1, p + 1):
                hh = h
                h = [0.0] * (j + 1)
                for i in range(1, j + 1):
                    tli, tlj = tk[lm + i], tk[lm + i - j]
                    if tli != tlj:
                        f = hh[i - 1] / (tli - tlj)
                        h[i - 1] += f * (tli - x)
                        h[i] = f * (x - tlj)
            coef[row, m] = sum(cm[lm - p + j] * h[j] for j in range(p + 1)) / math.factorial(m)
    return brk, coef


class HostTables:
    def __init__(self, path=_DATA):
        d = np.load(path)
        self.xbmin, self.xbmax = float(d["xbmin"]), float(d["xbmax"])
        self.xfmin, self.xfmax = float(d["xfmin"]), float(d["xfmax"])
        self.tb, self.cb, self.tf, self.cf = d["tb"], d["cb"], d["tf"], d["cf"]
        # the 2 x 1000 exact-integral samples the splines interpolate (finiteT.py:151-196)
        self.xb, self.yb, self.xf, self.yf = d["xb"], d["yb"], d["xf"], d["yf"]
        self.low_b, self.low_f = np.ascontiguousarray(d["low_b"]), np.ascontiguousarray(d["low_f"])
        self.brk_b, self.coef_b = bspline_to_pp(self.tb, self.cb)
        self.brk_f, self.coef_f = bspline_to_pp(self.tf, self.cf)
        # closed-form locator: the samples sit at x_k = |xmin| sinh(u_k)/20, u_k uniform
        # (finiteT.py:159-161,189-191) -> (scale, u0, 1/du) with u = asinh(x * scale)
        self.loc_b = self._locator(d["xb"], self.xbmin)
        self.loc_f = self._locator(d["xf"], self.xfmin)

    @staticmethod
    def _locator(x, xmin):
        scale = 20.0 / abs(xmin)
        u = np.arcsinh(x * scale)
        du = (u[-1] - u[0]) / (len(x) - 1)
        assert np.max(np.abs(np.diff(u) - du)) < 1e-9 * du, "sample grid is not sinh-uniform"
        return scale, float(u[0]), 1.0 / du


_host = None
_device = {}


def host_tables():
    global _host
    if _host is None:
        _host = HostTables()
    return _host


class DeviceTables:
    """Device copies of both tables + the ctb_jtable structs that point at them."""

    def __init__(self, device):
        torch = _lib.require_cuda()
        h = host_tables()
        self.device = device
        self._keep = []
        self.jb = self._make(torch, h.brk_b, h.coef_b, h.xbmin, h.xbmax, h.loc_b)
        self.jf = self._make(torch, h.brk_f, h.coef_f, h.xfmin, h.xfmax, h.loc_f)

    def _make(self, torch, brk, coef, xmin, xmax, loc):
        b = torch.from_numpy(np.ascontiguousarray(brk)).to(self.device)
        c = torch.from_numpy(np.ascontiguousarray(coef)).to(self.device)
        assert c.data_ptr() % 32 == 0
        self._keep += [b, c]
        return _lib.JTable(b.data_ptr(), c.data_ptr(), coef.shape[0], 0, xmin, xmax, *loc)


def device_tables(device):
    torch = _lib.require_cuda()
    device = torch.device(device)
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    key = device.index
    if key not in _device:
        _device[key] = DeviceTables(device)
    return _device[key]
