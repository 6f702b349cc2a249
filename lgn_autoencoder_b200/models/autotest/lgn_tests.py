"""Equivariance acceptance test (reference: lgn/models/autotest/lgn_tests.py:23-423): boosts and rotations about
an axis are applied to the input jets, and the autoencoder's outputs and all internal node features must
transform with the corresponding Lorentz-D matrices."""
import logging
import time
from math import cosh, sqrt

import numpy as np
import torch

from ...g_lib import rotations as rot
from .utils import display_err, get_avg_internal_dev, get_avg_output_dev, get_dev, get_output


def _gen_rot(angles, maxdim, device=torch.device("cpu"), dtype=torch.float64, cg_dict=None):
    """Lorentz-D matrices for all irreps below maxdim and the Cartesian 4x4 matrix R acting on input momenta as
    p' = p R (derived from D^(1,1) through the canonical <-> Cartesian basis change)."""
    D = {(k, n): rot.LorentzD((k, n), *angles, device=device, dtype=dtype, cg_dict=cg_dict) for k in range(maxdim) for n in range(maxdim)}
    r = 1 / sqrt(2.0)
    m = torch.tensor([[1, 0, 0, 0], [0, r, -1j * r, 0], [0, 0, 0, 1], [0, -r, -1j * r, 0]], dtype=torch.complex128, device=device)
    d11 = torch.complex(D[(1, 1)][0], D[(1, 1)][1]).to(torch.complex128)
    R = (m.conj().T @ d11 @ m).real.to(dtype)
    return D, R


def get_rotation(theta, axis):
    return {"x": (theta, 0, 0), "y": (0, theta, 0)}.get(axis.lower(), (0, 0, theta))


def get_boost(alpha, axis):
    return {"x": (alpha * 1j, 0, 0), "y": (0, alpha * 1j, 0)}.get(axis.lower(), (0, 0, alpha * 1j))


def _equivariance(encoder, decoder, data, params, angle_fn, axis, device, dtype, cg_dict):
    t_in, t_out, t_in_nodes, t_out_nodes = [], [], [], []
    res, internal = get_output(encoder, decoder, data, covariance_test=True)
    for a in params:
        angles = angle_fn(a, axis)
        _, R = _gen_rot(angles, encoder.maxdim, device=device, dtype=dtype, cg_dict=cg_dict)
        moved = dict(data)
        moved["p4"] = torch.einsum("...b,ba->...a", data["p4"], R)
        res_in, internal_in = get_output(encoder, decoder, moved, covariance_test=True)
        t_in.append(res_in)
        t_in_nodes.append(internal_in)
        t_out.append(rot.rotate_rep(res, *angles, cg_dict=cg_dict))
        t_out_nodes.append([rot.rotate_rep(x, *angles, cg_dict=cg_dict) for x in internal])
    return get_dev(t_in, t_out, t_in_nodes, t_out_nodes, mode="mean")


def boost_equivariance(encoder, decoder, data, alpha_range, axis, device, dtype, cg_dict):
    dev_output, dev_internal = _equivariance(encoder, decoder, data, alpha_range, get_boost, axis, device, dtype, cg_dict)
    return [cosh(x) for x in alpha_range], dev_output, dev_internal


def rot_equivariance(encoder, decoder, data, theta_range, axis, device, dtype, cg_dict):
    dev_output, dev_internal = _equivariance(encoder, decoder, data, theta_range, get_rotation, axis, device, dtype, cg_dict)
    return theta_range, dev_output, dev_internal


def covariance_test(encoder, decoder, data, test_type, axis="z", alpha_max=None, cg_dict=None, unit="GeV"):
    if cg_dict is None:
        cg_dict = encoder.cg_dict
    device, dtype = encoder.device, encoder.dtype
    data = dict(data)
    data["p4"] = data["p4"].to(device, dtype).clone()
    if unit.lower() == "gev":
        data["p4"] = data["p4"] / 1e3
    out = {}
    if test_type.lower() in ("boost", "boosts"):
        alpha_max = 10.0 if alpha_max is None else alpha_max
        alphas = np.arange(0, alpha_max + 0.01, step=alpha_max / 25.0)
        out["gammas"], out["boost_dev_output"], out["boost_dev_internal"] = boost_equivariance(encoder, decoder, data, alphas, axis, device, dtype, cg_dict)
    elif test_type.lower() in ("rot", "rotation", "rotations"):
        alpha_max = 2 * np.pi if alpha_max is None else alpha_max
        thetas = np.arange(0, alpha_max + 0.01, step=alpha_max / 25.0)
        out["thetas"], out["rot_dev_output"], out["rot_dev_internal"] = rot_equivariance(encoder, decoder, data, thetas, axis, device, dtype, cg_dict)
    else:
        raise ValueError(f"test_type must be one of 'boost' or 'rotation': {test_type}")
    return out


@torch.no_grad()
def permutation_invariance_test(encoder, decoder, data, *ignore):
    """Deviation of the reconstruction under a random permutation of the input particles."""
    device, dtype = encoder.device, encoder.dtype
    p4 = data["p4"].to(device, dtype)
    perm = torch.randperm(p4.shape[1], device=device)
    moved = dict(data)
    moved["p4"] = p4[:, perm]
    if "labels" in data:
        moved["labels"] = data["labels"].to(device)[:, perm]
    base = dict(data)
    base["p4"] = p4
    out = decoder(encoder(base, covariance_test=False), covariance_test=False)
    out_perm = decoder(encoder(moved, covariance_test=False), covariance_test=False)
    inv = (out - out_perm).abs().max().item() / (out.abs().max().item() + 1e-16)
    equi = (out[:, :, perm] - out_perm).abs().max().item() / (out.abs().max().item() + 1e-16)
    return {"invariance": inv, "equivariance": equi}


def lgn_tests(args, encoder, decoder, dataloader, alpha_max=None, theta_max=None, cg_dict=None, unit="GeV", axis="z"):
    """Run the boost / rotation / permutation tests over the batches of ``dataloader`` and return the averaged
    deviation tables ({'gammas', 'boost_dev_output', 'boost_dev_internal', 'thetas', 'rot_dev_output', ...})."""
    t0 = time.time()
    encoder.eval()
    decoder.eval()
    boosts, rots, perms = [], [], []
    for data in dataloader:
        boosts.append(covariance_test(encoder, decoder, data, "boost", axis=axis, alpha_max=alpha_max, cg_dict=cg_dict, unit=unit))
        rots.append(covariance_test(encoder, decoder, data, "rotation", axis=axis, alpha_max=theta_max, cg_dict=cg_dict, unit=unit))
        perms.append(permutation_invariance_test(encoder, decoder, data))
    results = {
        "gammas": boosts[0]["gammas"], "thetas": rots[0]["thetas"],
        "boost_dev_output": get_avg_output_dev(boosts, "boost"), "boost_dev_internal": get_avg_internal_dev(boosts, "boost"),
        "rot_dev_output": get_avg_output_dev(rots, "rot"), "rot_dev_internal": get_avg_internal_dev(rots, "rot"),
        "perm_invariance_dev_output": sum(p["invariance"] for p in perms) / len(perms),
        "perm_equivariance_dev_output": sum(p["equivariance"] for p in perms) / len(perms),
    }
    display_err(results["gammas"], results["boost_dev_output"], "gamma", "Boost equivariance deviation")
    display_err(results["thetas"], results["rot_dev_output"], "theta", "Rotation equivariance deviation")
    logging.info(f"Equivariance tests took {time.time() - t0:.1f} s")
    return results
