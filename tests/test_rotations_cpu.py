"""Host logic of the equivariance harness (lgn_autoencoder_b200/g_lib/rotations.py, reference lgn/g_lib/rotations.py:7-210) on
CPU: the Lorentz-D matrices are a representation, D^(1,1) is a Lorentz transformation of the Cartesian momenta, and the
Clebsch-Gordan tables of the oracle intertwine them -- CG(D u (x) D v) = D CG(u (x) v) -- for rotations and boosts.  These are
the mathematical known-answer tests SURVEY.md section 4 lists; no GPU, no golden file."""
import numpy as np
import pytest
import torch

from lgn_autoencoder_b200.cg_lib import CGDict
from lgn_autoencoder_b200.g_lib import rotations as rot
from oracle import lgae_oracle as orc

ANGLES = [("rot_z", (0.0, 0.0, 0.7), (0.0, 0.0, -1.9)), ("rot_y", (0.0, 0.4, 0.0), (0.0, 1.1, 0.0)),
          ("boost_z", (0.0, 0.0, 0.8j), (0.0, 0.0, 1.7j))]
KEYS = [(0, 0), (1, 1), (0, 2), (2, 0), (2, 2)]


@pytest.fixture(scope="module")
def cg():
    return CGDict(maxdim=3, dtype=torch.float64, device=torch.device("cpu"))


def dmat(key, angles, cg):
    return rot.LorentzD(key, *angles, numpy_test=True, cg_dict=cg)


@pytest.mark.parametrize("name,a,b", ANGLES)
def test_lorentz_d_is_a_representation(name, a, b, cg):
    """About a fixed axis the group is abelian: D(a) D(b) = D(a + b), D(0) = 1, D(-a) = D(a)^-1."""
    for key in KEYS:
        d = (key[0] + 1) * (key[1] + 1)
        da, db = dmat(key, a, cg), dmat(key, b, cg)
        dab = dmat(key, tuple(x + y for x, y in zip(a, b)), cg)
        assert np.allclose(da @ db, dab, atol=1e-12), (name, key)
        assert np.allclose(dmat(key, (0.0, 0.0, 0.0), cg), np.eye(d), atol=1e-13)
        assert np.allclose(dmat(key, tuple(-x for x in a), cg) @ da, np.eye(d), atol=1e-12), (name, key)


@pytest.mark.parametrize("name,a,b", ANGLES)
def test_d11_is_a_lorentz_transformation(name, a, b, cg):
    """The Cartesian matrix the harness applies to the input momenta (lgn_tests._gen_rot) preserves the Minkowski metric; a
    boost with rapidity alpha has gamma = cosh(alpha)."""
    from lgn_autoencoder_b200.models.autotest.lgn_tests import _gen_rot
    _, R = _gen_rot(a, 3, cg_dict=cg)
    eta = torch.diag(torch.tensor([1.0, -1.0, -1.0, -1.0], dtype=torch.float64))
    assert torch.allclose(R @ eta @ R.T, eta, atol=1e-12), name
    if name == "boost_z":
        assert abs(abs(R[0, 0].item()) - np.cosh(0.8)) < 1e-12


@pytest.mark.parametrize("name,a,b", ANGLES)
def test_cg_tables_intertwine_the_representations(name, a, b, cg):
    """CG(D1 u (x) D2 v) = D CG(u (x) v) for every pair of irreps below maxdim 3 and every output irrep: ties the oracle's CG
    coefficients (pinned to the reference's, test_oracle_golden) to the D matrices the equivariance harness rotates with, in
    the harness's own convention z -> z . conj(D) (rotate_part, side='left')."""
    table = orc.cg_table(3)
    g = torch.Generator().manual_seed(5)
    for (k1, k2), entry in table.items():
        d1, d2 = (k1[0] + 1) * (k1[1] + 1), (k2[0] + 1) * (k2[1] + 1)
        u = torch.complex(torch.randn(d1, generator=g, dtype=torch.float64), torch.randn(d1, generator=g, dtype=torch.float64)).numpy()
        v = torch.complex(torch.randn(d2, generator=g, dtype=torch.float64), torch.randn(d2, generator=g, dtype=torch.float64)).numpy()
        D1, D2 = np.conj(dmat(k1, a, cg)), np.conj(dmat(k2, a, cg))
        ur, vr = u @ D1, v @ D2                       # rotate_part(side='left'): z . conj(D)
        for ko, h in entry.items():
            if max(ko) > 2:      # cut by the model's maxdim (cg_ops.py:176-215); its D matrix would need a larger CG dictionary
                continue
            H = h.numpy()
            out = H @ np.kron(u, v)
            out_r = H @ np.kron(ur, vr)
            Do = np.conj(dmat(ko, a, cg))
            assert np.allclose(out_r, out @ Do, atol=1e-11), (name, k1, k2, ko, np.abs(out_r - out @ Do).max())
