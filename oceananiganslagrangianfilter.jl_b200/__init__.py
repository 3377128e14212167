"""lfb200 -- B200-native hot path of OceananigansLagrangianFilter.jl (offline filter PDEs).

The package directory is ``oceananiganslagrangianfilter.jl_b200/``; import it as ``lfb200``
through the shim module at the repository root.
"""
__version__ = "0.1.0"
