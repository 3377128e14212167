"""lfb200 -- B200-native hot path of OceananigansLagrangianFilter.jl (offline filter PDEs).

The package directory is ``oceananiganslagrangianfilter.jl_b200/``; import it as ``lfb200`` through
the shim module at the repository root.  The exported names mirror the reference's export list
(reference src/OceananigansLagrangianFilter.jl:23-30, src/Utils/Utils.jl:17-21); everything that
touches field data runs in liblfb200.so (hand-written CUDA for sm_100a) -- there is no CPU fallback.
"""
from .config import Centered, InputData, OfflineFilterConfig, OnlineFilterConfig, UpwindBiased, WENO
from .filter_utils import (FilterParams, create_filtered_vars, create_forcing, set_offline_BW2_filter_params,
                           set_online_BW_filter_params)
from .grids import (Bounded, Flat, GridFittedBottom, ImmersedBoundaryGrid, PartialCellBottom, Periodic,
                    RectilinearGrid)
from .jld2 import JLD2File
from .offline import aligned_time_steps, run_offline_Lagrangian_filter
from .post import (FilterOutput, compute_Eulerian_filter, compute_time_shift, get_offline_frequency_response,
                   get_online_frequency_response, get_weight_function, regrid_to_mean_position,
                   sum_forward_backward_contributions)

__version__ = "0.1.0"

__all__ = [
    "OfflineFilterConfig", "OnlineFilterConfig", "InputData", "WENO", "Centered", "UpwindBiased", "RectilinearGrid", "ImmersedBoundaryGrid", "GridFittedBottom",
    "PartialCellBottom", "Periodic", "Bounded", "Flat",
    "run_offline_Lagrangian_filter", "aligned_time_steps", "set_offline_BW2_filter_params",
    "set_online_BW_filter_params", "create_filtered_vars", "create_forcing", "FilterParams", "FilterOutput",
    "sum_forward_backward_contributions", "regrid_to_mean_position", "compute_Eulerian_filter", "compute_time_shift",
    "get_weight_function", "get_offline_frequency_response", "get_online_frequency_response", "JLD2File",
]


def LagrangianFilter(*args, **kwargs):
    """Device model (see model.LagrangianFilter); imported lazily so the host-only API works without the .so."""
    from .model import LagrangianFilter as _LF
    return _LF(*args, **kwargs)
