"""symboltz.jl_b200 -- B200-native (sm_100a) implementation of SymBoltz.jl's data-parallel hot path:
perturbation solve over independent k-modes -> source functions -> line-of-sight integration -> C_l, and P(k).
See DESIGN.md.  Import as `import symboltz.jl_b200 as sb` (the `symboltz` shim package at the repo root maps the
dotted name onto this directory)."""
from .api import (ΛCDM, LCDM, w0waCDM, ModelSpec, parameters_Planck18, CosmologyProblem, parameter_updater, solve, solvebg, solvebg_batch, solvept,
                  issuccess, spectrum_primordial, spectrum_matter, source_grid, source_grid_adaptive, refine_grid, source_kinterp, ChebyshevInterpolator, SourceGrid,
                  SphericalBesselCache, los_integrate, spectrum_cmb, spectrum_cmb_from_theta, natural_spline_weights, spline_ls, cmb_grids,
                  lingrid, loggrid, cosgrid, cospi, chebgrid, chebpoints, momentum_quadrature, k0, RETCODES, CMBPlan, spectrum_matter_sweep, fk_tanh, fk_tanh_inv, solvept_batch, cosmo_record, BatchSolution, COSMO_DTYPE, CosmoArena, spectrum_cmb_batch, sensitivity_matter, sensitivity_cmb, sensitivity_background, split_capacity, split_pays, ptalg,
                  build_schedule, ModeCostModel, resident_warps, shard_rows, gather_rows, solvebg_lock, solvept_lanes, source_background, SbmSrc, Communicator, CosmologySolution, CubicSplineInterpolator, EquispacedInterpolator, PiecewiseChebyshevInterpolator)
from . import build

__all__ = [n for n in dir() if not n.startswith("_")]
