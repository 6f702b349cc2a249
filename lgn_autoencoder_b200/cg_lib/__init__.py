"""Clebsch-Gordan machinery (API of the reference's lgn/cg_lib/__init__.py:1-23)."""
from .cg_dict import CGDict, sl2c_cg, su2_cg
from .cg_module import CGModule
from .cg_ops import CGProduct, cg_product, cg_product_tau
from .zonal_functions import (ZonalFunctions, ZonalFunctionsRel, eps, metric, normsq4, p_cplx_to_rep, p_to_rep, rep_to_p, repdot,
                              zonal_functions, zonal_functions4, zonal_functions_canonical, zonal_functions_rel)

__all__ = ["CGDict", "CGModule", "CGProduct", "cg_product", "cg_product_tau", "ZonalFunctions", "ZonalFunctionsRel", "normsq4",
           "p_to_rep", "p_cplx_to_rep", "rep_to_p", "repdot", "metric", "zonal_functions", "zonal_functions_rel", "eps"]
