"""CPU-only checks of the drop-in boundary: the shared library loads, exports every symbol include/lgae_b200.h declares,
and the argument-checking / geometry entry points behave without a GPU (no compute calls)."""
import ctypes as C
import os
import re
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "lgae_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lgae_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    from lgn_autoencoder_b200 import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert getattr(lib, n) is not None, n
    # the ctypes prototypes cover the whole header
    assert set(names) == set(_lib.EXPORTED_SYMBOLS), set(names) ^ set(_lib.EXPORTED_SYMBOLS)


def test_version_and_error_strings():
    from lgn_autoencoder_b200 import _lib
    lib = _lib.load()
    assert lib.lgae_version() >= 100
    assert lib.lgae_error_string(0) == b"ok"
    for code in (-1, -2, -3, -4):
        assert len(lib.lgae_error_string(code)) > 3


def _plan(n=30):
    from lgn_autoencoder_b200.fused import FusedPlan
    import torch
    from tests.helpers import load_golden
    g = load_golden("cfg1_b3")
    cfg = g["cfg"]
    return FusedPlan("encoder", OrderedDict((k, tuple(v.shape)) for k, v in g["enc_state"].items()), n_particles=n,
                     channels=cfg["enc_channels"], latent_mode=cfg["map_to_latent"], tau_s=cfg["tau_s"], tau_v=cfg["tau_v"],
                     num_basis_fn=10, mlp=True, mlp_depth=6, mlp_width=6), g


def test_geometry_entry_points_without_gpu():
    from lgn_autoencoder_b200 import _lib
    lib = _lib.load()
    plan, g = _plan()
    d = C.byref(plan.desc)
    assert plan.n_params == sum(v.numel() for v in g["enc_state"].values()) == 34146
    w1, w8 = lib.lgae_workspace_doubles(d, 1), lib.lgae_workspace_doubles(d, 8)
    assert 0 < w1 < w8 and (w8 - lib.lgae_workspace_doubles(d, 0)) > 0
    assert lib.lgae_workspace_doubles(d, -1) == -1
    # saved tensors sit inside the workspace, 32-byte aligned, in increasing order per kind
    offs = [lib.lgae_workspace_offset(d, 8, kind, 0) for kind in range(9)]
    assert all(0 <= o < w8 and o % 4 == 0 for o in offs), offs
    assert lib.lgae_workspace_offset(d, 8, 0, 99) == -1 and lib.lgae_workspace_offset(d, 8, 42, 0) == -1
    assert lib.lgae_partials_doubles(d, 8) > 0
    assert lib.lgae_mlp_pack_doubles(d, 0) > 0 and lib.lgae_mlp_pack_doubles(d, 7) == -1


def test_bad_arguments_are_rejected_before_any_launch():
    from lgn_autoencoder_b200 import _lib
    lib = _lib.load()
    plan, _ = _plan()
    d = C.byref(plan.desc)
    # null pointers -> LGAE_E_BADARG, never a crash; an encoder descriptor is refused by the decoder entry point
    assert lib.lgae_encoder_forward(d, None, None, None, 4, None, None, None, None, None) == -1
    assert lib.lgae_decoder_forward(d, None, None, 4, None, None, None, None) == -1
    assert lib.lgae_chamfer(None, None, 4, 30, 30, 2, None, None, None, None, None) == -1
    assert lib.lgae_chamfer(None, None, 0, 30, 30, 7, None, None, None, None, None) == -1   # unknown get_real mode
    assert lib.lgae_normalize_p4(None, 4, 30, None, None, None) == -1
    assert lib.lgae_encoder_forward(None, None, None, None, 4, None, None, None, None, None) == -1
    bad = _lib.LgaeModelDesc()
    assert lib.lgae_workspace_doubles(C.byref(bad), 4) == -1
    # the saved-radial-weights buffer of the level entry point is a 32-particle layout: refused (LGAE_E_UNSUPPORTED) for longer
    # jets before anything is launched (the pointers are never dereferenced on the host)
    plan40, _ = _plan(40)
    fake = C.c_void_p(0x1000)
    assert lib.lgae_level_forward(C.byref(plan40.desc), 0, fake, fake, None, 2, fake, fake, fake, fake, fake, fake, None) == -2
    assert lib.lgae_level_forward(C.byref(plan40.desc), 9, fake, fake, None, 2, fake, fake, fake, fake, fake, fake, None) == -1


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    """No CPU fallback: without the CUDA library the product import path raises."""
    from lgn_autoencoder_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    try:
        _lib.load()
    except ImportError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("load() must fail when the library is missing")
    # monkeypatch restores the handle and the path (no reload: that would re-create the ctypes classes the handle is bound to)


def test_product_never_imports_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "lgn_autoencoder_b200")):
        for f in files:
            if f.endswith(".py") and re.search(r"^\s*(from|import)\s+oracle", open(os.path.join(dirpath, f)).read(), flags=re.M):
                bad.append(f)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "lgn")):
        for f in files:
            if f.endswith(".py") and re.search(r"^\s*(from|import)\s+oracle", open(os.path.join(dirpath, f)).read(), flags=re.M):
                bad.append(f)
    assert not bad, bad


def test_layer_level_entry_points_check_arguments_without_gpu():
    """lgae_cg_product_* / lgae_mix_* / lgae_scalar_irrep_* / lgae_radial_functions_* / lgae_linear_*: size queries and
    argument validation return before any CUDA call."""
    from lgn_autoencoder_b200 import _lib
    lib = _lib.load()
    BADARG = -1
    assert lib.lgae_mix_partials_doubles(-1, 4, 4) == -1 and lib.lgae_mix_partials_doubles(100, 0, 4) == -1
    assert lib.lgae_mix_partials_doubles(10, 3, 5) == 10 * 2 * 3 * 5
    assert lib.lgae_linear_partials_doubles(1000, 8, 0) == -1 and lib.lgae_linear_partials_doubles(100, 8, 6) == 6 * 9
    assert lib.lgae_radial_functions_partials_doubles(-5, 20, 8, 2) == -1
    rows = lib.lgae_radial_functions_partials_doubles(128, 20, 8, 2)   # whole rows of 2*8*21 + 60 partials, one per CTA
    assert rows >= 2 * 8 * 21 + 60 and rows % (2 * 8 * 21 + 60) == 0
    assert lib.lgae_cg_product_forward(None, None, None, None, None, 4, 0, None, None) == BADARG
    d = _lib.LgaeCgPairDesc()
    d.d1, d.d2, d.channels, d.n_out, d.n_comp, d.n_terms = 4, 4, 3, 17, 1, 1          # more output irreps than LGAE_CG_MAX_OUT
    assert lib.lgae_cg_product_forward(C.byref(d), 1, 1, None, None, 4, 0, 1, None) == BADARG
    assert lib.lgae_mix_forward(None, None, 5, 0, 3, 4, None, None) == BADARG
    assert lib.lgae_mix_forward(None, None, 0, 2, 3, 4, None, None) == 0                # empty input: nothing to launch
    assert lib.lgae_scalar_irrep_forward(None, None, 7, 3, 2, 4, None, None) == BADARG  # channel counts do not broadcast
    assert lib.lgae_scalar_irrep_forward(None, None, 0, 3, 1, 4, None, None) == 0
    assert lib.lgae_linear_forward(None, None, None, 0, 6, 36, 1, 0.01, None, None) == 0
    assert lib.lgae_linear_forward(None, None, None, 5, 6, 36, 1, 0.01, None, None) == BADARG
    assert lib.lgae_radial_functions_forward(None, None, 0, 1, None, None, None, 20, 8, 2, None, None, None, 1, None) == 0


def test_adam_entry_point_checks_arguments_without_gpu():
    from lgn_autoencoder_b200 import _lib
    lib = _lib.load()
    assert lib.lgae_adam_step(None, None, None, None, 0, None, None, None, None, 0, 1e-3, 0.9, 0.999, 1e-8, 0.0, None, None) == -1   # no step state
    assert lib.lgae_adam_step(None, None, None, None, 0, None, None, None, None, 0, 1e-3, 0.9, 0.999, 1e-8, 0.0, 8, None) == 0      # nothing to update
    assert lib.lgae_adam_step(None, None, None, None, 5, None, None, None, None, 0, 1e-3, 0.9, 0.999, 1e-8, 0.0, 8, None) == -1     # NULL buffers
    assert lib.lgae_adam_step(None, None, None, None, 0, None, None, None, None, 0, 1e-3, 1.0, 0.999, 1e-8, 0.0, 8, None) == -1     # beta1 == 1
