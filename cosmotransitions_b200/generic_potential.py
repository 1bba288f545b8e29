"""GPU evaluation of generic_potential.Vtot = V0 + V1 + V1T (generic_potential.py:240-337) for the
model of examples/testModel1.py (V0: :62-80, boson_massSq: :82-116; fermions: the base-class
placeholder with zero degrees of freedom, generic_potential.py:226-238)."""
import ctypes as C

import numpy as np

from . import _lib, _tables
from .potentials import model1_model8, Model1LinePotential


class model1:
    """Same constructor parameters as examples/testModel1.py:model1.init."""

    Ndim = 2

    def __init__(self, m1=120., m2=50., mu=25., Y1=.1, Y2=.15, n=30):
        self.model8 = model1_model8(m1, m2, mu, Y1, Y2, n)
        self.l1, self.l2, self.mu2, self.Y1, self.Y2, self.n, self.renormScaleSq, self.v2 = self.model8
        self._m8 = (C.c_double * 8)(*self.model8)

    def Vtot(self, X, T, include_radiation=True):
        """Total finite-temperature effective potential; X[..., 2] and T broadcast as in the
        reference (generic_potential.py:318-324).  model1 sets no num_boson_dof, so
        include_radiation changes nothing (generic_potential.py:289-295)."""
        return self._eval(X, T, _lib.V_TOTAL)

    def V1T_from_X(self, X, T, include_radiation=True):
        """One-loop finite-temperature part alone (generic_potential.py:298-310)."""
        return self._eval(X, T, _lib.V_THERMAL)

    def DVtot(self, X, T):
        """Vtot offset such that V(0, T) = 0 (generic_potential.py:339-345)."""
        torch = _lib.require_cuda()
        X0 = np.zeros(self.Ndim)
        return self._eval(X, T, _lib.V_TOTAL) - self._eval(X0, T, _lib.V_TOTAL)

    def _eval(self, X, T, parts):
        torch = _lib.require_cuda()
        lib = _lib.load()
        is_tensor = isinstance(X, torch.Tensor)
        dev = X.device if is_tensor and X.is_cuda else torch.device("cuda", torch.cuda.current_device())
        Xd = (X if is_tensor else torch.from_numpy(np.asarray(X, dtype=np.float64))).to(dev, torch.float64)
        Td = (T if isinstance(T, torch.Tensor) else torch.from_numpy(np.asarray(T, dtype=np.float64))).to(
            dev, torch.float64)
        shape = torch.broadcast_shapes(Xd.shape[:-1], Td.shape)
        Xf = Xd.expand(shape + (2,)).contiguous().reshape(-1, 2)
        Tf = Td.expand(shape).contiguous().reshape(-1)
        out = torch.empty(Tf.shape, dtype=torch.float64, device=dev)
        tabs = _tables.device_tables(dev)
        with torch.cuda.device(dev):
            _lib.check(lib.ctb_vtot_model1(self._m8, tabs.jb, Xf.data_ptr(), Tf.data_ptr(), out.data_ptr(), Tf.numel(),
                                           parts, torch.cuda.current_stream(dev).cuda_stream), "ctb_vtot_model1")
        out = out.reshape(shape)
        if is_tensor:
            return out if X.is_cuda else out.cpu()
        res = out.cpu().numpy()
        return float(res) if res.ndim == 0 else res

    def Vtot_grid(self, x0, nhat, s, T, out=None):
        """Vtot on the outer-product grid X_i = x0 + s_i nhat, T_j -> [len(s), len(T)] (device tensor
        if s is one).  One pass, output only: 8 B of HBM traffic per point."""
        torch = _lib.require_cuda()
        lib = _lib.load()
        is_tensor = isinstance(s, torch.Tensor)
        dev = s.device if is_tensor and s.is_cuda else torch.device("cuda", torch.cuda.current_device())
        sd = (s if is_tensor else torch.from_numpy(np.asarray(s, dtype=np.float64))).to(dev, torch.float64).contiguous()
        Td = (T if isinstance(T, torch.Tensor) else torch.from_numpy(np.asarray(T, dtype=np.float64))).to(
            dev, torch.float64).contiguous()
        if out is None:
            out = torch.empty((sd.numel(), Td.numel()), dtype=torch.float64, device=dev)
        line = (C.c_double * 4)(float(x0[0]), float(x0[1]), float(nhat[0]), float(nhat[1]))
        tabs = _tables.device_tables(dev)
        with torch.cuda.device(dev):
            _lib.check(lib.ctb_vtot_model1_grid(self._m8, line, tabs.jb, sd.data_ptr(), sd.numel(), Td.data_ptr(),
                                                Td.numel(), out.data_ptr(),
                                                torch.cuda.current_stream(dev).cuda_stream), "ctb_vtot_model1_grid")
        if is_tensor and s.is_cuda:
            return out
        return out.cpu() if is_tensor else out.cpu().numpy()

    def line_potential(self, x_abs, x_meta, T):
        """Device functor of Vtot along the straight line from x_abs (phi = 0) to x_meta (phi = L)."""
        x_abs, x_meta = np.asarray(x_abs, float), np.asarray(x_meta, float)
        L = float(np.linalg.norm(x_meta - x_abs))
        return Model1LinePotential(x_abs, (x_meta - x_abs) / L, T, self.model8), L
