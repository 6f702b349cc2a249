"""Lorentz-group representation matrices used by the equivariance acceptance test
(reference: lgn/g_lib/rotations.py:7-210).  Host-side numpy/torch; nothing here is on the hot path.

Conventions: multiplet components run m = -j..+j ascending; d^j(beta) = exp(i beta J_y)
with J_y = -(J+ - J-)/(2i) built on the descending-m ladder the reference uses; D^j(a,b,c) = e^{i a m} d^j(b) e^{i c m'};
D^{(k,n)} = C [D^{k/2}(a,b,c) (x) conj D^{n/2}(-a,b,-c)] C^T with C the ((k,0),(0,n)) -> (k,n) Clebsch-Gordan matrix.
Boosts are imaginary angles."""
import numpy as np
import torch


def _jy(j):
    m = -np.arange(-j, j)
    ladder = np.sqrt((j + m) * (j - m + 1))
    jp, jm = np.diag(ladder, k=1), np.diag(ladder, k=-1)
    return -(jp - jm) / 2j


def littled(j, beta):
    evals, evecs = np.linalg.eigh(_jy(j))
    return evecs @ np.diag(np.exp(1j * beta * evals)) @ evecs.conj().T


def wigner_d_numpy(j, alpha, beta, gamma):
    m = np.arange(-j, j + 1)
    return np.exp(1j * alpha * m)[:, None] * littled(j, beta) * np.exp(1j * gamma * m)[None, :]


def complex_from_numpy(z, dtype=torch.float64, device=torch.device("cpu")):
    z = np.asarray(z, dtype=np.complex128)
    return torch.stack((torch.from_numpy(z.real.copy()), torch.from_numpy(z.imag.copy())), 0).to(dtype=dtype, device=device)


def WignerD(j, alpha, beta, gamma, numpy_test=False, dtype=torch.float64, device=torch.device("cpu")):
    d = wigner_d_numpy(j, alpha, beta, gamma)
    return d if numpy_test else complex_from_numpy(d, dtype=dtype, device=device)


def LorentzD(key, alpha, beta, gamma, numpy_test=False, dtype=torch.float64, device=torch.device("cpu"), cg_dict=None):
    k, n = key
    if cg_dict is None:
        from ..cg_lib import CGDict
        cg_dict = CGDict(maxdim=max(k, n) + 1, transpose=True, dtype=dtype, device=device)
    cg = cg_dict[((k, 0), (0, n))][(k, n)].detach().cpu().numpy().astype(np.float64)
    d1 = wigner_d_numpy(k / 2, alpha, beta, gamma)
    d2 = np.conj(wigner_d_numpy(n / 2, -alpha, beta, -gamma))
    big = cg @ np.kron(d1, d2) @ cg.T
    return big if numpy_test else complex_from_numpy(big, dtype=dtype, device=device)


def dagger(D):
    return torch.stack((D[0], -D[1]), 0).transpose(-1, -2)


def conj(D):
    return torch.stack((D[0], -D[1]), 0)


def rotate_part(D, z, side="left", autoconvert=True, conjugate=False):
    """side='left': z . conj(D) on the representation axis; side='right': conj(D) . z."""
    if autoconvert:
        D = D.to(z.device, z.dtype)
    if conjugate:
        D = dagger(D)
    Dc = torch.complex(D[0], -D[1])
    zc = torch.complex(z[0], z[1])
    out = torch.matmul(zc, Dc) if side == "left" else torch.matmul(Dc, zc)
    if side not in ("left", "right"):
        raise ValueError("Must choose side: left/right.")
    return torch.stack((out.real, out.imag), 0)


def rotate_rep(rep, alpha, beta, gamma, side="left", conjugate=False, cg_dict=None):
    device, dtype = rep.device, rep.dtype
    return rep.__class__({key: rotate_part(LorentzD(key, alpha, beta, gamma, cg_dict=cg_dict, device=device, dtype=dtype), part,
                                           side=side, conjugate=conjugate) for key, part in rep.items()})
