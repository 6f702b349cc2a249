"""Equivariance acceptance test (reference: lgn/models/autotest/lgn_tests.py:23-423): boosts and rotations about
an axis are applied to the input jets, and the autoencoder's outputs and all internal node features must
transform with the corresponding Lorentz-D matrices."""
import logging
import time
from math import cosh, sqrt

import numpy as np
import torch

from ...g_lib import rotations as rot
from .utils import display_err, get_avg_internal_dev, get_avg_output_dev, get_dev, get_node_dev, get_output


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


MAX_JETS_PER_PASS = 16384   # transforms are batched into one model evaluation of up to this many jets


def _slice_rep(rep, lo, hi):
    return rep.__class__({k: v[:, lo:hi] for k, v in rep.items()})


def _equivariance(encoder, decoder, data, params, angle_fn, axis, device, dtype, cg_dict):
    """f(T x) against T f(x) for every transform T of `params`.  The reference evaluates the model twice per transform
    (lgn_tests.py:179-269: 26 boosts + 26 rotations -> 104 evaluations per batch); jets are independent, so here ALL
    transformed copies of the batch go through the autoencoder as one batch of len(params) * B jets (per-slice R), and the
    untransformed batch is evaluated once."""
    res, internal = get_output(encoder, decoder, data, covariance_test=True)
    p4 = data["p4"]
    B = p4.shape[0]
    angles_all = [angle_fn(a, axis) for a in params]
    Rs = torch.stack([_gen_rot(ang, encoder.maxdim, device=device, dtype=dtype, cg_dict=cg_dict)[1] for ang in angles_all])   # (T,4,4)
    chunk = max(1, MAX_JETS_PER_PASS // max(B, 1))
    t_in, t_out, t_in_nodes, t_out_nodes = [], [], [], []
    for t0 in range(0, len(angles_all), chunk):
        R = Rs[t0:t0 + chunk]
        T = R.shape[0]
        moved = {k: (torch.cat([v] * T, 0) if torch.is_tensor(v) and v.dim() >= 1 and v.shape[0] == B else v) for k, v in data.items()}
        moved["p4"] = torch.einsum("bnm,tma->tbna", p4, R).reshape(T * B, *p4.shape[1:])
        res_in, internal_in = get_output(encoder, decoder, moved, covariance_test=True)
        for t in range(T):
            angles = angles_all[t0 + t]
            t_in.append(_slice_rep(res_in, t * B, (t + 1) * B))
            t_in_nodes.append([_slice_rep(x, t * B, (t + 1) * B) for x in internal_in])
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
    """lgn_tests.py:140-176: the real particles of every jet are permuted (padding stays in place); returns the 'max'-mode
    deviations (invariance: f(P x) vs f(x); equivariance: f(P x) vs P f(x)) of the generated (0,0) and (1,1) features.
    The two evaluations run as one batch of 2 B jets."""
    device, dtype = encoder.device, encoder.dtype
    p4 = data["p4"].to(device, dtype)
    mask = data["labels"].to(device) if "labels" in data else (p4[..., 0] != 0).to(torch.uint8)
    B, N = mask.shape
    perm = torch.arange(N).expand(B, -1).clone()
    for idx in range(B):
        n_real = int(mask[idx].long().sum())
        perm[idx, :n_real] = torch.randperm(n_real)
    perm = perm.to(device)

    def apply_perm(mat):
        return torch.gather(mat, 1, perm.view(B, N, *([1] * (mat.dim() - 2))).expand_as(mat))

    assert (mask == apply_perm(mask)).all()
    both = {k: (torch.cat([v.to(device), v.to(device)], 0) if torch.is_tensor(v) and v.dim() >= 1 and v.shape[0] == B else v) for k, v in data.items()}
    both["p4"] = torch.cat([p4, apply_perm(p4)], 0)
    if "scalars" in data:
        sc = data["scalars"].to(device)
        both["scalars"] = torch.cat([sc, apply_perm(sc)], 0)
    out, _ = get_output(encoder, decoder, both, covariance_test=True)
    noperm = {k: v[:, :B].squeeze() for k, v in out.items()}
    permed = {k: v[:, B:].squeeze() for k, v in out.items()}
    perm_outputs = {k: torch.stack((apply_perm(v[0]), apply_perm(v[1])), 0) for k, v in noperm.items()}
    return get_node_dev(permed, noperm, mode="max"), get_node_dev(permed, perm_outputs, mode="max")


@torch.no_grad()
def lgn_tests(args, encoder, decoder, dataloader, axis="z", alpha_max=None, theta_max=None, cg_dict=None, unit="GeV"):
    """lgn_tests.py:292-423 (same positional order): boost / rotation / permutation tests over the first
    ``args.num_test_batch`` batches of ``dataloader``; returns the averaged deviation tables ({'gammas', 'boost_dev_output',
    'boost_dev_internal', 'thetas', 'rot_dev_output', 'rot_dev_internal', 'perm_invariance_dev_output',
    'perm_equivariance_dev_output'})."""
    t0 = time.time()
    encoder.eval()
    decoder.eval()
    boosts, rots, perm_inv, perm_equi = [], [], [], []
    num_test_batch = getattr(args, "num_test_batch", -1)
    for idx, data in enumerate(dataloader):
        boosts.append(covariance_test(encoder, decoder, data, "boost", axis=axis, alpha_max=alpha_max, cg_dict=cg_dict, unit=unit))
        rots.append(covariance_test(encoder, decoder, data, "rotation", axis=axis, alpha_max=theta_max, cg_dict=cg_dict, unit=unit))
        inv, equi = permutation_invariance_test(encoder, decoder, data)
        perm_inv.append(inv)
        perm_equi.append(equi)
        if num_test_batch is not None and num_test_batch > 0 and idx + 1 >= num_test_batch:
            break
    print(f"Covariance test completed! Time taken: {round((time.time() - t0) / 60, 2)} min")
    results = {
        "gammas": boosts[0]["gammas"], "thetas": rots[0]["thetas"],
        "boost_dev_output": get_avg_output_dev(boosts, "boost"), "boost_dev_internal": get_avg_internal_dev(boosts, "boost"),
        "rot_dev_output": get_avg_output_dev(rots, "rot"), "rot_dev_internal": get_avg_internal_dev(rots, "rot"),
        "perm_invariance_dev_output": {k: sum(d[k] for d in perm_inv) / len(perm_inv) for k in perm_inv[0]},
        "perm_equivariance_dev_output": {k: sum(d[k] for d in perm_equi) / len(perm_equi) for k in perm_equi[0]},
    }
    print(display_err(results["gammas"], results["boost_dev_output"], "gamma", "Boost equivariance test result: output relative error"))
    print(display_err(results["thetas"], results["rot_dev_output"], "theta", "Rotation equivariance test result: output relative error"))
    print(f"Permutation invariance test result: {results['perm_invariance_dev_output']}")
    print(f"Permutation equivariance test result: {results['perm_equivariance_dev_output']}")
    return results
