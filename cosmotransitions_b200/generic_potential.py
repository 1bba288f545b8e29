"""GPU evaluation of generic_potential.Vtot = V0 + V1 + V1T (generic_potential.py:240-337) for the
model of examples/testModel1.py (V0: :62-80, boson_massSq: :82-116; fermions: the base-class
placeholder with zero degrees of freedom, generic_potential.py:226-238)."""
import ctypes as C

import numpy as np

from . import _lib, _tables
from .potentials import model1_model8, Model1LinePotential, Thermal1FPotential, thermal1f_params


def _stencil1(eps, order, deriv):
    """(offset, weight) pairs of helper_functions.gradientFunction / hessianFunction for one direction
    (helper_functions.py:446-461, 522-528): first or second derivative, order 2 or 4."""
    if order == 2:
        return ([(-eps, -.5 / eps), (eps, .5 / eps)] if deriv == 1 else
                [(-eps, 1. / (eps * eps)), (0.0, -2. / (eps * eps)), (eps, 1. / (eps * eps))])
    if deriv == 1:
        return [(k * eps, w / (12. * eps)) for k, w in ((-2, 1.), (-1, -8.), (1, 8.), (2, -1.))]
    return [(k * eps, w / (12. * eps * eps)) for k, w in ((-2, -1.), (-1, 16.), (0, -30.), (1, 16.), (2, -1.))]


class _FiniteDifferences:
    """gradV, d2V, dgradV_dT, energyDensity, massSqMatrix of generic_potential (generic_potential.py:347-456) as finite
    differences of the device Vtot / V1T: every stencil point of every input point goes into ONE kernel launch
    (`self._combo`), the weighted sum is formed on the device.  Needs Ndim, x_eps, T_eps, deriv_order, _combo."""

    x_eps, T_eps, deriv_order = .001, .001, 4

    def _unit(self, i, h):
        e = np.zeros(self.Ndim)
        e[i] = h
        return e

    def gradV(self, X, T):
        """Gradient of Vtot (radiation terms off); same shape as X (generic_potential.py:347-367)."""
        comps = [self._combo(X, T, _lib.V_TOTAL, False, [(self._unit(i, h), 0.0, w)
                                                        for h, w in _stencil1(self.x_eps, self.deriv_order, 1)])
                 for i in range(self.Ndim)]
        return _stack(comps, -1)

    def d2V(self, X, T, parts=None):
        """Hessian of Vtot (radiation terms off); shape X.shape + (Ndim,) (generic_potential.py:419-442)."""
        parts = _lib.V_TOTAL if parts is None else parts
        s1, s2 = _stencil1(self.x_eps, self.deriv_order, 1), _stencil1(self.x_eps, self.deriv_order, 2)
        rows = []
        for i in range(self.Ndim):
            row = []
            for j in range(self.Ndim):
                if i == j:
                    sh = [(self._unit(i, h), 0.0, w) for h, w in s2]
                else:   # product of two first-derivative stencils (helper_functions.py:505-519)
                    sh = [(self._unit(i, hi) + self._unit(j, hj), 0.0, wi * wj) for hi, wi in s1 for hj, wj in s1]
                row.append(self._combo(X, T, parts, False, sh) if j >= i else None)
            rows.append(row)
        for i in range(self.Ndim):
            for j in range(i):
                rows[i][j] = rows[j][i]
        return _stack([_stack(r, -1) for r in rows], -2)

    def massSqMatrix(self, X):
        """Hessian of the tree-level potential V0 (generic_potential.py:390-417)."""
        return self.d2V(X, 0.0, parts=_lib.V_TREE)

    def dgradV_dT(self, X, T):
        """d/dT of the gradient, from V1T alone (generic_potential.py:369-388)."""
        sx, sT = _stencil1(self.x_eps, self.deriv_order, 1), _stencil1(self.T_eps, self.deriv_order, 1)
        comps = [self._combo(X, T, _lib.V_THERMAL, False, [(self._unit(i, hx), hT, wx * wT) for hx, wx in sx for hT, wT in sT])
                 for i in range(self.Ndim)]
        return _stack(comps, -1)

    def energyDensity(self, X, T, include_radiation=True):
        """V - T dV/dT with the temperature derivative taken from V1T (generic_potential.py:443-456)."""
        dVdT = self._combo(X, T, _lib.V_THERMAL, include_radiation,
                           [(np.zeros(self.Ndim), hT, w) for hT, w in _stencil1(self.T_eps, self.deriv_order, 1)])
        V = self._combo(X, T, _lib.V_TOTAL, include_radiation, [(np.zeros(self.Ndim), 0.0, 1.0)])
        Tb = T if hasattr(dVdT, "is_cuda") else np.asarray(T, dtype=np.float64)
        return V - Tb * dVdT


def _stack(parts, axis):
    if hasattr(parts[0], "is_cuda"):
        import torch
        return torch.stack(list(parts), dim=axis)
    return np.stack([np.asarray(p) for p in parts], axis=axis)


class model1(_FiniteDifferences):
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
        return self._combo(X, T, parts, True, None)

    def _combo(self, X, T, parts, include_radiation, shifts):
        """sum_k w_k V(X + dx_k, T + dT_k) over shifts = [(dx[2], dT, w), ...] in one launch (None: plain evaluation).
        model1 sets no num_boson_dof / num_fermion_dof, so include_radiation changes nothing."""
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
        n = Tf.numel()
        if shifts is not None:
            Xf = torch.cat([Xf + torch.as_tensor(dx, dtype=torch.float64, device=dev) for dx, _, _ in shifts]).contiguous()
            Tf = torch.cat([Tf + dT for _, dT, _ in shifts]).contiguous()
        if Xf.data_ptr() % 16:      # the kernel reads one 16-byte (phi1, phi2) pair per point
            Xf = Xf.clone()
        out = torch.empty(Tf.shape, dtype=torch.float64, device=dev)
        tabs = _tables.device_tables(dev)
        with torch.cuda.device(dev):
            _lib.check(lib.ctb_vtot_model1(self._m8, tabs.jb, Xf.data_ptr(), Tf.data_ptr(), out.data_ptr(), Tf.numel(),
                                           parts, torch.cuda.current_stream(dev).cuda_stream), "ctb_vtot_model1")
        if shifts is not None:
            out = sum(w * out[k * n:(k + 1) * n] for k, (_, _, w) in enumerate(shifts))
        out = out.reshape(shape)
        if is_tensor:
            return out if X.is_cuda else out.cpu()
        res = out.cpu().numpy()
        return float(res) if res.ndim == 0 else res

    def Vtot_grid(self, x0, nhat, s, T, out=None, host_out=None):
        """Vtot on the outer-product grid X_i = x0 + s_i nhat, T_j -> [len(s), len(T)] (device tensor
        if s is one).  One pass, output only: 8 B of HBM traffic per point.  host_out: an optional page-locked CPU
        tensor of that shape that receives the grid for host callers (the 33.5 MB of a 4096 x 1024 grid take 15 ms through
        pageable memory and 1.4 ms through a pinned buffer); the returned numpy array is then a view of it."""
        torch = _lib.require_cuda()
        lib = _lib.load()
        is_tensor = isinstance(s, torch.Tensor)
        dev = s.device if is_tensor and s.is_cuda else torch.device("cuda", torch.cuda.current_device())
        sd = (s if is_tensor else torch.from_numpy(np.asarray(s, dtype=np.float64))).to(dev, torch.float64).contiguous()
        Td = (T if isinstance(T, torch.Tensor) else torch.from_numpy(np.asarray(T, dtype=np.float64))).to(
            dev, torch.float64).contiguous()
        if out is None:
            out = torch.empty((sd.numel(), Td.numel()), dtype=torch.float64, device=dev)
        elif not (isinstance(out, torch.Tensor) and out.device == dev and out.dtype == torch.float64
                  and out.is_contiguous() and tuple(out.shape) == (sd.numel(), Td.numel())):
            raise ValueError("out must be a contiguous float64 tensor of shape (len(s), len(T)) on %s" % (dev,))
        line = (C.c_double * 4)(float(x0[0]), float(x0[1]), float(nhat[0]), float(nhat[1]))
        tabs = _tables.device_tables(dev)
        with torch.cuda.device(dev):
            _lib.check(lib.ctb_vtot_model1_grid(self._m8, line, tabs.jb, sd.data_ptr(), sd.numel(), Td.data_ptr(),
                                                Td.numel(), out.data_ptr(),
                                                torch.cuda.current_stream(dev).cuda_stream), "ctb_vtot_model1_grid")
        if is_tensor and s.is_cuda:
            return out
        if host_out is not None:
            if not (isinstance(host_out, torch.Tensor) and not host_out.is_cuda and host_out.dtype == torch.float64
                    and host_out.is_contiguous() and tuple(host_out.shape) == tuple(out.shape)):
                raise ValueError("host_out must be a contiguous float64 CPU tensor of shape (len(s), len(T))")
            host_out.copy_(out, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
            return host_out if is_tensor else host_out.numpy()
        return out.cpu() if is_tensor else out.cpu().numpy()

    def findMinimum(self, X=None, T=0.0, max_iter=60, xtol=1e-9):
        """generic_potential.findMinimum (generic_potential.py:477-484: one scipy.optimize.fmin call per point) for any
        number of points at once, entirely on the device: X [..., 2] against T [...] -> minima of the broadcast shape +
        (2,).  One launch of the warp-per-point damped Newton kernel (ctb_trace_minima, csrc/ctb_phases.cuh); points
        whose iteration does not end on a minimum come back as NaN."""
        from . import phases
        if X is None:
            X = self.approxZeroTMin()[0]
        X = np.asarray(X, dtype=np.float64)
        shape = np.broadcast_shapes(X.shape[:-1], np.shape(T))
        Xb = np.ascontiguousarray(np.broadcast_to(X, shape + (2,))).reshape(-1, 2)
        Tb = np.ascontiguousarray(np.broadcast_to(np.asarray(T, dtype=np.float64), shape)).reshape(-1)
        r = phases.find_minimum(self.model8, Xb, Tb, max_iter=max_iter, xtol=xtol, x_eps=self.x_eps)
        out = r["X"].cpu().numpy()
        out[r["flag"].cpu().numpy() != _lib.MIN_OK] = np.nan
        return out.reshape(shape + (2,))

    def approxZeroTMin(self):
        """The two tree-level minima (examples/testModel1.py:118-122)."""
        v = float(self.v2) ** .5
        return [np.array([v, v]), np.array([v, -v])]

    def line_potential(self, x_abs, x_meta, T):
        """Device functor of Vtot along the straight line from x_abs (phi = 0) to x_meta (phi = L)."""
        x_abs, x_meta = np.asarray(x_abs, float), np.asarray(x_meta, float)
        L = float(np.linalg.norm(x_meta - x_abs))
        return Model1LinePotential(x_abs, (x_meta - x_abs) / L, T, self.model8), L


class one_field_model(_FiniteDifferences):
    """A general ONE-field generic_potential (generic_potential.py:26-345) on the device: quartic tree-level
    potential V0 = sum_k V0_coefs[k] phi^k, bosons with m^2 = a + b phi^2 + t T^2 (dof, c), fermions with
    m^2 = a + b phi^2 (dof) -- the data a subclass would return from V0 / boson_massSq / fermion_massSq
    (generic_potential.py:132-238), given once as tables instead of Python callbacks.

    Same attributes (Ndim, x_eps, T_eps, deriv_order, renormScaleSq, num_boson_dof, num_fermion_dof) and the same
    methods with the reference's broadcasting contract (X[..., 1] against T[...]): Vtot, V1T_from_X, DVtot, gradV,
    d2V, dgradV_dT, energyDensity.  `potential(T)` hands the same model to the bounce solver."""

    Ndim = 1

    def __init__(self, V0_coefs, renormScaleSq=1000. ** 2, bosons=(), fermions=(), num_boson_dof=None,
                 num_fermion_dof=None, x_eps=.001, T_eps=.001, deriv_order=4):
        self.V0_coefs, self.bosons, self.fermions = list(V0_coefs), [tuple(b) for b in bosons], [tuple(f) for f in fermions]
        self.renormScaleSq, self.x_eps, self.T_eps, self.deriv_order = renormScaleSq, x_eps, T_eps, deriv_order
        self.num_boson_dof, self.num_fermion_dof = num_boson_dof, num_fermion_dof
        if deriv_order not in (2, 4):
            raise ValueError("deriv_order must be 2 or 4")
        row = thermal1f_params(self.V0_coefs, renormScaleSq, 0.0, self.bosons, self.fermions)[0]
        self._row = (C.c_double * len(row))(*row)

    def potential(self, T):
        """Device functor of Vtot(phi, T) for SingleFieldInstanton / find_bounce_batch."""
        return Thermal1FPotential(self.V0_coefs, self.renormScaleSq, T, self.bosons, self.fermions)

    def _rad(self, include_radiation):
        rad = 0.0   # generic_potential.py:289-295
        if include_radiation and self.num_boson_dof is not None:
            rad += (self.num_boson_dof - sum(b[3] for b in self.bosons)) * np.pi ** 4 / 45.
        if include_radiation and self.num_fermion_dof is not None:
            rad += (self.num_fermion_dof - sum(f[2] for f in self.fermions)) * 7 * np.pi ** 4 / 360.
        return rad

    def _combo(self, X, T, parts, include_radiation=True, shifts=None):
        """sum_k w_k V(X + dx_k, T + dT_k) for shifts = [(dx[1], dT, w), ...] in one launch (finite differences)."""
        torch = _lib.require_cuda()
        lib = _lib.load()
        is_tensor = isinstance(X, torch.Tensor)
        dev = X.device if is_tensor and X.is_cuda else torch.device("cuda", torch.cuda.current_device())
        Xd = (X if is_tensor else torch.from_numpy(np.asarray(X, dtype=np.float64))).to(dev, torch.float64)
        Td = (T if isinstance(T, torch.Tensor) else torch.from_numpy(np.asarray(T, dtype=np.float64))).to(
            dev, torch.float64)
        if Xd.dim() == 0 or Xd.shape[-1] != 1:
            raise ValueError("X must have a last axis of length Ndim = 1")
        shape = torch.broadcast_shapes(Xd.shape[:-1], Td.shape)
        phi = Xd[..., 0].expand(shape).reshape(-1)
        Tf = Td.expand(shape).reshape(-1)
        shifts = shifts or [(np.zeros(1), 0.0, 1.0)]
        phis = torch.cat([phi + float(np.asarray(dx).reshape(-1)[0]) for dx, _, _ in shifts]).contiguous()
        Ts = torch.cat([Tf + dT for _, dT, _ in shifts]).contiguous()
        out = torch.empty_like(phis)
        tabs = _tables.device_tables(dev)
        with torch.cuda.device(dev):
            _lib.check(lib.ctb_vtot_thermal1f(self._row, self._rad(include_radiation), tabs.jb, tabs.jf, phis.data_ptr(),
                                              Ts.data_ptr(), out.data_ptr(), phis.numel(), parts,
                                              torch.cuda.current_stream(dev).cuda_stream), "ctb_vtot_thermal1f")
        n = phi.numel()
        res = sum(w * out[k * n:(k + 1) * n] for k, (_, _, w) in enumerate(shifts)).reshape(shape)
        if is_tensor:
            return res if X.is_cuda else res.cpu()
        res = res.cpu().numpy()
        return float(res) if res.ndim == 0 else res

    def findMinimum(self, X, T=0.0, max_iter=60, xtol=1e-9):
        """generic_potential.findMinimum (generic_potential.py:477-484) for any number of points, entirely on the device
        (ctb_trace_minima with the one-field model: warp-per-point damped Newton on the reference's finite-difference
        derivatives).  X [..., 1] against T [...]; points whose iteration does not end on a minimum come back as NaN."""
        from . import phases
        X = np.asarray(X, dtype=np.float64)
        shape = np.broadcast_shapes(X.shape[:-1], np.shape(T))
        Xb = np.ascontiguousarray(np.broadcast_to(X, shape + (1,))).reshape(-1, 1)
        Tb = np.ascontiguousarray(np.broadcast_to(np.asarray(T, dtype=np.float64), shape)).reshape(-1)
        r = phases.find_minimum(np.array(self._row), Xb, Tb, max_iter=max_iter, xtol=xtol, x_eps=self.x_eps, model="thermal1f")
        out = r["X"].cpu().numpy()
        out[r["flag"].cpu().numpy() != _lib.MIN_OK] = np.nan
        return out.reshape(shape + (1,))

    def Vtot(self, X, T, include_radiation=True):
        """The total finite temperature effective potential (generic_potential.py:312-337)."""
        return self._combo(X, T, _lib.V_TOTAL, include_radiation)

    def V1T_from_X(self, X, T, include_radiation=True):
        """The one-loop finite-temperature part alone (generic_potential.py:298-310)."""
        return self._combo(X, T, _lib.V_THERMAL, include_radiation)

    def DVtot(self, X, T):
        """Vtot offset such that V(0, T) = 0 (generic_potential.py:339-345)."""
        X0 = np.zeros(self.Ndim)
        return self.Vtot(X, T, False) - self.Vtot(X0, T, False)


class compiled_model(_FiniteDifferences):
    """A generic_potential whose Vtot(X, T) is a run-time compiled CUDA body (jit.CompiledModel): the counterpart of
    subclassing the reference's generic_potential with new V0 / boson_massSq / fermion_massSq
    (generic_potential.py:132-238) for ANY number of fields up to four.  `params` is this model's parameter row.

    Same broadcasting contract as the reference (X[..., Ndim] against T[...]): Vtot, DVtot, gradV, d2V (the reference's
    finite-difference stencils, every stencil point in one launch), findMinimum (device Newton), plus
    trace_minima / phase_scan / line_bounces through `phases` with `model=self.model`."""

    def __init__(self, model, params, x_eps=.001, T_eps=.001, deriv_order=4):
        self.model, self.Ndim = model, model.ndim
        self.params = np.ascontiguousarray(params, dtype=np.float64).reshape(max(1, model.nparam))
        self.x_eps, self.T_eps, self.deriv_order = x_eps, T_eps, deriv_order

    def _combo(self, X, T, parts, include_radiation=True, shifts=None):
        if parts != _lib.V_TOTAL:
            raise NotImplementedError("a compiled model evaluates Vtot as a whole (V0 / V1 / V1T are not separable)")
        torch = _lib.require_cuda()
        is_tensor = isinstance(X, torch.Tensor)
        dev = X.device if is_tensor and X.is_cuda else torch.device("cuda", torch.cuda.current_device())
        Xd = (X if is_tensor else torch.from_numpy(np.asarray(X, dtype=np.float64))).to(dev, torch.float64)
        Td = (T if isinstance(T, torch.Tensor) else torch.from_numpy(np.asarray(T, dtype=np.float64))).to(dev, torch.float64)
        if Xd.dim() == 0 or Xd.shape[-1] != self.Ndim:
            raise ValueError("X must have a last axis of length Ndim = %d" % self.Ndim)
        shape = torch.broadcast_shapes(Xd.shape[:-1], Td.shape)
        Xf = Xd.expand(shape + (self.Ndim,)).reshape(-1, self.Ndim)
        Tf = Td.expand(shape).reshape(-1)
        n = Tf.numel()
        shifts = shifts or [(np.zeros(self.Ndim), 0.0, 1.0)]
        Xs = torch.cat([Xf + torch.as_tensor(np.asarray(dx, dtype=np.float64), device=dev) for dx, _, _ in shifts]).contiguous()
        Ts = torch.cat([Tf + dT for _, dT, _ in shifts]).contiguous()
        out = self.model.vtot(self.params, Xs, Ts)
        res = sum(w * out[k * n:(k + 1) * n] for k, (_, _, w) in enumerate(shifts)).reshape(shape)
        if is_tensor:
            return res if X.is_cuda else res.cpu()
        res = res.cpu().numpy()
        return float(res) if res.ndim == 0 else res

    def Vtot(self, X, T, include_radiation=True):
        return self._combo(X, T, _lib.V_TOTAL)

    def DVtot(self, X, T):
        return self.Vtot(X, T) - self.Vtot(np.zeros(self.Ndim), T)

    def findMinimum(self, X, T=0.0, max_iter=60, xtol=1e-9):
        """generic_potential.findMinimum (generic_potential.py:477-484) for any number of points, on the device."""
        from . import phases
        X = np.asarray(X, dtype=np.float64)
        shape = np.broadcast_shapes(X.shape[:-1], np.shape(T))
        Xb = np.ascontiguousarray(np.broadcast_to(X, shape + (self.Ndim,))).reshape(-1, self.Ndim)
        Tb = np.ascontiguousarray(np.broadcast_to(np.asarray(T, dtype=np.float64), shape)).reshape(-1)
        r = phases.find_minimum(self.params, Xb, Tb, max_iter=max_iter, xtol=xtol, x_eps=self.x_eps, model=self.model)
        out = r["X"].cpu().numpy()
        out[r["flag"].cpu().numpy() != _lib.MIN_OK] = np.nan
        return out.reshape(shape + (self.Ndim,))
