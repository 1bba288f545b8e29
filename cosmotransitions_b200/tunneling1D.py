"""Drop-in for cosmoTransitions.tunneling1D.SingleFieldInstanton (tunneling1D.py:49-826) whose
numerical work -- __init__ (barrier location, rscale), findProfile, findAction -- runs on the GPU as a
batch of one through the same kernel the batched entry point uses.

Same constructor signature, same findProfile knobs and defaults, same Profile1D namedtuple, same
exceptions (PotentialError(msg, kind) with kind in {"no barrier", "stable, not metastable"},
helper_functions.IntegrationError), so it can be handed to pathDeformation.fullTunneling as
`tunneling_class=` (pathDeformation.py:849,959-968) for potentials the device knows.

V is a device functor (cosmotransitions_b200.potentials.DevicePotential): dV / d2V are then taken from the
functor (analytic where it has them, otherwise the reference's own 5-point stencils, tunneling1D.py:151-161,191-201,
evaluated on the device) and the solve is trajectory-faithful to the reference.  Plain Python callables -- which the
reference accepts, tunneling1D.py:112-149 -- cannot run inside a CUDA kernel and there is no CPU fallback for the
integration: they are sampled once onto the tabulated functor (potentials.SampledPotential: `V_spline_pieces`
cubic pieces between the two minima, Hermite data from dV when it is given) and the solve runs on the GPU on that
table -- spline-accurate (S converges like pieces^-4), not trajectory-exact.  SplinePath.V / .dV / .d2V of the
reference's path deformation are recognised and use the path's own spline exactly.
"""
from collections import namedtuple

import numpy as np

from . import _lib, batch
from .helper_functions import IntegrationError, monotonicIndices
from .potentials import DevicePotential


class PotentialError(Exception):
    """
    Used when the potential does not have the expected characteristics.

    The error messages should be tuples, with the second item being one of
    ``("no barrier", "stable, not metastable")``.
    """
    pass


_ERRORS = {
    _lib.ST_STABLE: lambda: PotentialError("V(phi_metaMin) <= V(phi_absMin); tunneling cannot occur.",
                                           "stable, not metastable"),
    _lib.ST_NO_BARRIER: lambda: PotentialError("The potential barrier does not exist between phi_bar and "
                                               "phi_metaMin.", "no barrier"),
    _lib.ST_RMAX: lambda: IntegrationError("r > rmax"),
    _lib.ST_DRMIN: lambda: IntegrationError("dr < drmin"),
    _lib.ST_STEP_UNDERFLOW: lambda: IntegrationError("Stepsize rounds down to zero."),
    _lib.ST_MAXSTEPS: lambda: IntegrationError("more than 2**20 integration steps (guard of the GPU solver)"),
    _lib.ST_NONFINITE: lambda: ValueError("The function value in the initialConditions root find is NaN; "
                                          "solver cannot continue."),
    _lib.ST_BRENT_SIGN: lambda: ValueError("f(a) and f(b) must have different signs"),
    _lib.ST_FIRST_SHOT: lambda: AssertionError("Failed to retrieve initial conditions on the first try."),
}


def _as_device_potential(V, dV, phi_absMin, phi_metaMin, pieces):
    """DevicePotential as is; SplinePath.V of the reference (pathDeformation.py:959-964: V, dV, d2V are splev() of one
    cubic spline) -> that spline, exactly; any other callable -> sampled (potentials.SampledPotential)."""
    if isinstance(V, DevicePotential):
        return V
    from .potentials import SampledPotential, SplinePotential
    if not callable(V):
        raise TypeError("V must be a DevicePotential or a callable")
    return SplinePotential.from_callable(V) or SampledPotential.for_instanton(V, dV, float(phi_absMin), float(phi_metaMin),
                                                                             pieces=pieces)


def raise_for_status(status):
    """Maps a per-instance status code to the exception the reference raises at that point."""
    status = int(status)
    if status in _ERRORS:
        raise _ERRORS[status]()
    if status not in (_lib.ST_CONVERGED, _lib.ST_XTOL):
        raise RuntimeError("bounce solver: " + _lib.status_string(status))


class SingleFieldInstanton:
    """GPU-backed counterpart of tunneling1D.SingleFieldInstanton (see module docstring)."""

    profile_rval = namedtuple("Profile1D", "R Phi dPhi Rerr")

    def __init__(self, phi_absMin, phi_metaMin, V, dV=None, d2V=None, phi_eps=1e-3, alpha=2, phi_bar=None,
                 rscale=None, V_spline_pieces=None):
        V = _as_device_potential(V, dV, phi_absMin, phi_metaMin, V_spline_pieces)
        self.phi_absMin, self.phi_metaMin = float(phi_absMin), float(phi_metaMin)
        self.V = V
        self.alpha = alpha
        self._phi_eps_rel = phi_eps
        self.phi_eps = phi_eps * abs(self.phi_absMin - self.phi_metaMin)
        self._last = None
        self._phi_bar_in, self._rscale_in = phi_bar, rscale
        # the reference validates the potential in __init__ (tunneling1D.py:133-147): do the same
        st, self.phi_bar, self.rscale = batch.bounce_setup(V.kind, V.params[None, :], [self.phi_absMin],
                                                          [self.phi_metaMin], phi_bar=phi_bar, rscale=rscale)
        if st in (_lib.ST_STABLE, _lib.ST_NO_BARRIER):
            raise_for_status(st)

    def _solve(self, return_profiles=False, **knobs):
        return batch.find_bounce_batch(self.V.kind, self.V.params[None, :], [self.phi_absMin], [self.phi_metaMin],
                                       alpha=self.alpha, phi_eps=self._phi_eps_rel, return_profiles=return_profiles,
                                       phi_bar=None if self._phi_bar_in is None else [self._phi_bar_in],
                                       rscale=None if self._rscale_in is None else [self._rscale_in], **knobs)

    def findProfile(self, xguess=None, xtol=1e-4, phitol=1e-4, thinCutoff=.01, npoints=500, rmin=1e-4, rmax=1e4,
                    max_interior_pts=None):
        r = self._solve(return_profiles=True, xguess=xguess, xtol=xtol, phitol=phitol, thinCutoff=thinCutoff,
                        npoints=npoints, rmin=rmin, rmax=rmax, max_interior_pts=max_interior_pts)
        raise_for_status(r["status"][0])
        n = int(r["n_points"][0])
        prof = self.profile_rval(r["R"][0, :n].numpy().copy(), r["Phi"][0, :n].numpy().copy(),
                                 r["dPhi"][0, :n].numpy().copy(), None)
        self._last = (prof, float(r["S"][0]))
        return prof

    def findAction(self, profile):
        """Euclidean action of `profile` (tunneling1D.py:763-791).  For the profile findProfile just
        returned this is the value the kernel accumulated while producing it (fused quadrature); any
        other Profile1D (anything with .R, .Phi, .dPhi) goes through the stand-alone quadrature kernel."""
        if self._last is not None and profile is self._last[0]:
            return self._last[1]
        S = batch.find_action_batch(self.V.kind, self.V.params[None, :], [self.phi_metaMin], np.asarray(profile.R, float),
                                    np.asarray(profile.Phi, float), np.asarray(profile.dPhi, float), alpha=self.alpha)
        return float(S[0])

    def evenlySpacedPhi(self, phi, dphi, npoints=100, k=1, fixAbs=True):
        """
        Re-grids `dphi` onto linearly spaced `phi` (tunneling1D.py:793-826; host-side post-processing of
        one ~750-point profile for the multi-field caller, pathDeformation.py:967-968).  k=1 (the only
        value the reference passes) is plain linear interpolation.
        """
        # pad with the end points the profile is known to reach (phi' = 0 there), drop whatever is not monotonic
        pad_phi, pad_dphi = ([self.phi_absMin], [0.0]) if fixAbs is True else ([], [])
        grid = np.concatenate([pad_phi, np.asarray(phi, dtype=float), [self.phi_metaMin]])
        vals = np.concatenate([pad_dphi, np.asarray(dphi, dtype=float), [0.0]])
        keep = monotonicIndices(grid)
        grid, vals = grid[keep], vals[keep]
        p = np.linspace(self.phi_absMin if fixAbs else grid[0], self.phi_metaMin, npoints)
        if k == 1:
            # a degree-1 interpolating spline is the piecewise-linear interpolant (np.interp wants ascending abscissae)
            dp = np.interp(p, grid, vals) if grid[0] <= grid[-1] else np.interp(p, grid[::-1], vals[::-1])
        else:
            from scipy import interpolate
            dp = interpolate.splev(p, interpolate.splrep(grid, vals, k=k))
        return p, dp


class WallWithConstFriction(SingleFieldInstanton):
    """GPU-backed counterpart of tunneling1D.WallWithConstFriction (tunneling1D.py:829-999): the bubble-wall
    profile with a constant friction term, phi'' + F phi' = V'(phi), found by over/undershooting in F.
    Same constructor as SingleFieldInstanton; rscale comes from the quadratic fit of the overridden findRScale
    (tunneling1D.py:838-864); findProfile returns Profile1D(R, Phi, dPhi, F, Rerr)."""

    profile_rval = namedtuple("Profile1D", "R Phi dPhi F Rerr")

    def __init__(self, phi_absMin, phi_metaMin, V, dV=None, d2V=None, phi_eps=1e-3, alpha=2, phi_bar=None,
                 rscale=None, V_spline_pieces=None):
        V = _as_device_potential(V, dV, phi_absMin, phi_metaMin, V_spline_pieces)
        if phi_bar is not None or rscale is not None:
            raise NotImplementedError("phi_bar / rscale overrides are not wired for WallWithConstFriction")
        self.phi_absMin, self.phi_metaMin = float(phi_absMin), float(phi_metaMin)
        self.V = V
        self.alpha = alpha   # unused by the wall's equation of motion; findAction (inherited) still reads it
        self._phi_eps_rel = phi_eps
        self.phi_eps = phi_eps * abs(self.phi_absMin - self.phi_metaMin)
        self._last = None
        # __init__ of the reference validates the potential (tunneling1D.py:133-147): a two-point profile is the
        # cheapest call that runs exactly that code on the device
        r = batch.find_wall_batch(V.kind, V.params[None, :], [self.phi_absMin], [self.phi_metaMin], phi_eps=phi_eps,
                                  npoints=2)
        self.phi_bar, self.rscale = float(r["phi_bar"][0]), float(r["rscale"][0])
        if int(r["status"][0]) in (_lib.ST_STABLE, _lib.ST_NO_BARRIER):
            if int(r["status"][0]) == _lib.ST_NO_BARRIER:
                raise PotentialError("Cannot fit the potential to a negative quadratic.", "no barrier")
            raise_for_status(r["status"][0])

    def findProfile(self, Fguess=None, Ftol=1e-4, phitol=1e-4, npoints=500, rmin=1e-4, rmax=1e4, phi0_rel=1e-3):
        r = batch.find_wall_batch(self.V.kind, self.V.params[None, :], [self.phi_absMin], [self.phi_metaMin],
                                  phi_eps=self._phi_eps_rel, return_profiles=True, Fguess=Fguess, Ftol=Ftol, phitol=phitol,
                                  npoints=npoints, rmin=rmin, rmax=rmax, phi0_rel=phi0_rel)
        raise_for_status(r["status"][0])
        n = int(r["n_points"][0])
        return self.profile_rval(r["R"][0, :n].numpy().copy(), r["Phi"][0, :n].numpy().copy(),
                                 r["dPhi"][0, :n].numpy().copy(), float(r["F"][0]), None)
