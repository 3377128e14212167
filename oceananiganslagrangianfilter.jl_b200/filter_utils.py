"""Host-side mirror of the filter-specification helpers in reference
src/Utils/lagrangian_filter_utils.jl: coefficients, tracer naming, forcing description.

These are pure host logic: they decide WHAT the kernels compute; the
arithmetic on field data is done by liblfb200.so.  Names, argument meaning and error behaviour follow
the reference so its tests translate line by line (tests/test_host_api.py).
"""
from __future__ import annotations

import cmath
import math
from typing import Dict, List, Tuple


class FilterParams(dict):
    """Dict with attribute access, standing in for the Julia NamedTuple `filter_params`
    (keys a1,b1,c1,d1,...,N_coeffs)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    @property
    def single_exponential(self) -> bool:
        return self["N_coeffs"] == 0.5

    @property
    def n_sets(self) -> int:
        return 1 if self.single_exponential else int(self["N_coeffs"])

    def coefficient_arrays(self):
        """(a, b, c, d) lists, with b = d = 0 for the single exponential."""
        n = self.n_sets
        a = [float(self[f"a{i}"]) for i in range(1, n + 1)]
        c = [float(self[f"c{i}"]) for i in range(1, n + 1)]
        if self.single_exponential:
            return a, [0.0] * n, c, [0.0] * n
        b = [float(self[f"b{i}"]) for i in range(1, n + 1)]
        d = [float(self[f"d{i}"]) for i in range(1, n + 1)]
        return a, b, c, d


def set_offline_BW2_filter_params(N: int = 1, freq_c: float = 1) -> FilterParams:
    """Butterworth-squared coefficients (reference utils:251-283)."""
    if N == 1:
        N_coeffs = 0.5
    elif N / 2 != math.floor(N / 2) or N < 0:
        raise ValueError("N must be a non-negative even integer, or 1 for a single exponential.")
    else:
        N_coeffs = int(N / 2)
    freq_c = abs(freq_c)
    if N_coeffs == 0.5:
        return FilterParams(a1=freq_c / 2, c1=freq_c, N_coeffs=N_coeffs)
    p = FilterParams()
    for i in range(1, N_coeffs + 1):
        th = math.pi / 2 / N * (2 * i - 1)
        p[f"a{i}"] = (freq_c / N) * math.sin(th)
        p[f"b{i}"] = (freq_c / N) * math.cos(th)
        p[f"c{i}"] = freq_c * math.sin(th)
        p[f"d{i}"] = freq_c * math.cos(th)
    p["N_coeffs"] = N_coeffs
    return p


def set_online_BW_filter_params(N: int = 1, freq_c: float = 1) -> FilterParams:
    """One-sided Butterworth coefficients (reference utils:321-361)."""
    if N == 1:
        N_coeffs = 0.5
    elif N / 2 != math.floor(N / 2) or N < 0:
        raise ValueError("N must be a non-negative even integer, or 1 for a single exponential.")
    else:
        N_coeffs = int(N / 2)
    freq_c = abs(freq_c)
    if N_coeffs == 0.5:
        return FilterParams(a1=freq_c, c1=freq_c, N_coeffs=N_coeffs)
    p = FilterParams()
    for i in range(1, N_coeffs + 1):
        c = freq_c * math.sin(math.pi / 2 / N * (2 * i - 1))
        d = -freq_c * math.cos(math.pi / 2 / N * (2 * i - 1))
        ri = cmath.exp(math.pi * 1j * (2 * i - 1) / (2 * N))
        Ai = 1
        for k in range(1, N + 1):
            if k != i:
                rk = cmath.exp(math.pi * 1j * (2 * k - 1) / (2 * N))
                Ai *= 1 / (ri - rk)
        rot = Ai * cmath.exp(-1j * math.pi * N / 2)
        p[f"a{i}"] = -2 * freq_c * rot.imag
        p[f"b{i}"] = 2 * freq_c * rot.real
        p[f"c{i}"] = c
        p[f"d{i}"] = d
    p["N_coeffs"] = N_coeffs
    return p


def validate_filter_params(filter_params) -> FilterParams:
    """The checks of OfflineLagrangianFilter.jl:371-405 on user-supplied coefficients."""
    p = FilterParams(filter_params)
    if "N_coeffs" in p:
        nc = p["N_coeffs"]
        if nc == 0.5:
            if not ("a1" in p and "c1" in p):
                raise ValueError("For N_coeffs=0.5, filter_params must have fields :a1, and :c1")
        elif int(math.floor(nc)) != nc:
            raise ValueError("N_coeffs must be a positive integer or 0.5")
        else:
            p["N_coeffs"] = int(nc)
            if not all(f"{k}{i}" in p for k in "abcd" for i in range(1, int(nc) + 1)):
                raise ValueError("For N_coeffs>0.5, filter_params must have fields :a1, :a2, ..., :b1, :b2, ..., "
                                 ":c1, :c2, ..., :d1, :d2, ...")
    else:
        n = len(p)
        if n % 4 == 0 and n > 0:
            p["N_coeffs"] = n // 4
            if not all(f"{k}{i}" in p for k in "abcd" for i in range(1, n // 4 + 1)):
                raise ValueError("filter_params must have fields :a1, :a2, ..., :b1, :b2, ..., :c1, :c2, ..., :d1, :d2, ...")
        elif n == 2:
            p["N_coeffs"] = 0.5
            if not ("a1" in p and "c1" in p):
                raise ValueError("For a filter with two coefficients, filter_params must have fields :a1, and :c1")
        else:
            raise ValueError("filter_params must have either 2 entries (for single exponential) or a multiple of 4 "
                             "entries, e.g. 2*N entries for Butterworth squared of order N, N even.")
    return p


def normalisation(filter_params: FilterParams) -> float:
    """sum (a c + b d)/(c^2 + d^2), which must be 1/2 for an offline filter
    (OfflineLagrangianFilter.jl:408-422); 2 a1 / c1 / 2 for the single exponential."""
    a, b, c, d = filter_params.coefficient_arrays()
    return sum((ai * ci + bi * di) / (ci * ci + di * di) for ai, bi, ci, di in zip(a, b, c, d))


# ---------------------------------------------------------------------------------------------- names
def create_filtered_vars(config) -> Tuple[str, ...]:
    """Tracer names in the reference's order (utils:423-470): all `_C` tracers (variables x
    coefficient sets, then `xi_<vel>` x coefficient sets), then all `_S` tracers."""
    fp = config.filter_params
    label = config.label
    gC: List[str] = []
    gS: List[str] = []
    with_maps = config.map_to_mean or config.compute_mean_velocities
    if fp["N_coeffs"] == 0.5:
        for v in config.var_names_to_filter:
            gC.append(f"{v}{label}_C1")
        if with_maps:
            for vel in config.velocity_names:
                gC.append(f"xi_{vel}{label}_C1")
        return tuple(gC)
    nc = int(fp["N_coeffs"])
    for v in config.var_names_to_filter:
        for i in range(1, nc + 1):
            gC.append(f"{v}{label}_C{i}")
            gS.append(f"{v}{label}_S{i}")
    if with_maps:
        for vel in config.velocity_names:
            for i in range(1, nc + 1):
                gC.append(f"xi_{vel}{label}_C{i}")
                gS.append(f"xi_{vel}{label}_S{i}")
    return tuple(gC) + tuple(gS)


def create_forcing(filtered_vars, config) -> Dict[str, tuple]:
    """Description of the forcing terms of every filter tracer (reference create_forcing, utils:731-865).

    The reference returns Oceananigans `Forcing` closures; here the terms are compiled into the fused
    stage kernel, so this returns, per tracer name, a tuple of term descriptors with the same
    multiplicity the reference's tuples have (its test checks those lengths,
    test/test_initialisation.jl:137-145)."""
    fp = config.filter_params
    label = config.label
    single = fp["N_coeffs"] == 0.5
    nc = 1 if single else int(fp["N_coeffs"])
    relax = bool(getattr(config, "boundary_relaxation", False))
    out: Dict[str, tuple] = {}
    n_var_tracers = (1 if single else 2) * nc * len(config.var_names_to_filter)
    for v in config.var_names_to_filter:
        for i in range(1, nc + 1):
            c = fp[f"c{i}"]
            d = 0.0 if single else fp[f"d{i}"]
            gC = (("filter_C", c, d), ("original_var", v))
            gS = (("filter_S", c, d),)
            if relax:
                gC += (("relaxation_C", v, config.relax_timescale),)
                gS += (("relaxation_S", v, config.relax_timescale),)
            out[f"{v}{label}_C{i}"] = gC
            if not single:
                # the reference stores a bare Forcing (not a tuple) when there is no relaxation (utils:822)
                out[f"{v}{label}_S{i}"] = gS
    if len(filtered_vars) > n_var_tracers:
        for vel in config.velocity_names:
            for i in range(1, nc + 1):
                c = fp[f"c{i}"]
                d = 0.0 if single else fp[f"d{i}"]
                xC = (("velocity_C", vel, c, d), ("filter_C", c, d))
                xS = (("velocity_S", vel, c, d), ("filter_S", c, d))
                if relax:
                    xC += (("relaxation_xiC", vel, config.relax_timescale),)
                    xS += (("relaxation_xiS", vel, config.relax_timescale),)
                out[f"xi_{vel}{label}_C{i}"] = xC
                if not single:
                    out[f"xi_{vel}{label}_S{i}"] = xS
    return out


def tracer_layout(config):
    """Map reference tracer name -> (field, pole, sine) in liblfb200's indexing: fields are the
    variables in `var_names_to_filter` order followed by one map per velocity in u, v, w order."""
    fp = config.filter_params
    label = config.label
    single = fp["N_coeffs"] == 0.5
    nc = 1 if single else int(fp["N_coeffs"])
    out = {}
    for q, v in enumerate(config.var_names_to_filter):
        for i in range(nc):
            out[f"{v}{label}_C{i + 1}"] = (q, i, 0)
            if not single:
                out[f"{v}{label}_S{i + 1}"] = (q, i, 1)
    if config.map_to_mean or config.compute_mean_velocities:
        vels = [v for v in ("u", "v", "w") if v in config.velocity_names]
        for m, vel in enumerate(vels):
            for i in range(nc):
                out[f"xi_{vel}{label}_C{i + 1}"] = (len(config.var_names_to_filter) + m, i, 0)
                if not single:
                    out[f"xi_{vel}{label}_S{i + 1}"] = (len(config.var_names_to_filter) + m, i, 1)
    return out
