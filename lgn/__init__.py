"""Drop-in ``lgn`` package: the reference's module API (lgn/models, lgn/nn, lgn/cg_lib, lgn/g_lib) served by
lgn_autoencoder_b200.  The reference's own ``lgn/`` is a namespace package without __init__.py, so putting this
repository on PYTHONPATH makes the unchanged main.py / test.py / covariance_test.py import this one
(SURVEY.md section 8(b)).  Sub-modules are aliased in sys.modules so that ``from lgn.g_lib import rotations`` and
``import lgn.models.autotest.lgn_tests`` resolve."""
import importlib
import sys

_PKG = "lgn_autoencoder_b200"
_SUBMODULES = [
    "g_lib", "g_lib.g_tau", "g_lib.g_tensor", "g_lib.cplx_lib", "g_lib.g_torch", "g_lib.rotations", "g_lib.weight_dict",
    "cg_lib", "cg_lib.cg_dict", "cg_lib.cg_module", "cg_lib.cg_ops", "cg_lib.zonal_functions",
    "nn", "nn.g_nn", "nn.position_levels", "nn.generic_levels",
    "models", "models.utils", "models.lgn_levels", "models.lgn_cg", "models.lgn_encoder", "models.lgn_decoder",
    "models.autotest", "models.autotest.lgn_tests", "models.autotest.utils",
]
for _name in _SUBMODULES:
    _mod = importlib.import_module(f"{_PKG}.{_name}")
    sys.modules[f"{__name__}.{_name}"] = _mod
    if "." not in _name:
        globals()[_name] = _mod
# names the reference exposes under different module paths
sys.modules[f"{__name__}.g_lib.g_vec"] = sys.modules[f"{__name__}.g_lib.g_tensor"]
sys.modules[f"{__name__}.g_lib.g_scalar"] = sys.modules[f"{__name__}.g_lib.g_tensor"]
sys.modules[f"{__name__}.g_lib.g_weight"] = sys.modules[f"{__name__}.g_lib.g_tensor"]
sys.modules[f"{__name__}.g_lib.parameter_dict_new"] = sys.modules[f"{__name__}.g_lib.weight_dict"]
sys.modules[f"{__name__}.cg_lib.cg_ops_tau"] = sys.modules[f"{__name__}.cg_lib.cg_ops"]
