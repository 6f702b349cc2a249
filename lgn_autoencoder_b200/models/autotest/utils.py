"""Deviation statistics of the equivariance harness (reference: lgn/models/autotest/utils.py:11-210).  The
reference's matplotlib plotting helpers are replaced by no-ops that keep the call signatures."""
import logging

import torch


@torch.no_grad()
def get_output(encoder, decoder, data, covariance_test=True):
    latent, enc_nodes = encoder(data, covariance_test=covariance_test)
    return decoder(latent, covariance_test=covariance_test, nodes_all=enc_nodes)


def get_node_dev(transform_input, transform_output, eps=1e-16, mode="mean"):
    """Relative deviation between f(T x) and T f(x) for the (0,0) and (1,1) parts.  'mean' is the metric of
    arXiv:2006.04780 used by the reference: |mean(a - b) / (mean(b) + eps)|."""
    keys = [(0, 0), (1, 1)]
    if mode.lower() == "max":
        return {w: ((transform_input[w] - transform_output[w]) / (transform_output[w] + eps)).abs().max().item() for w in keys}
    if mode.lower() != "mean":
        logging.warning(f"Mode {mode} not recognized. Returning mean.")
    return {w: abs((transform_input[w] - transform_output[w]).mean().item() / (transform_output[w].mean().item() + eps)) for w in keys}


def get_dev(transform_input, transform_output, transform_input_nodes_all, transform_output_nodes_all, mode="mean"):
    dev_output = [get_node_dev(a, b, mode=mode) for a, b in zip(transform_input, transform_output)]
    dev_internal = [[get_node_dev(a, b, mode=mode) for a, b in zip(ins, outs)]
                    for ins, outs in zip(transform_input_nodes_all, transform_output_nodes_all)]
    return dev_output, dev_internal


def _avg(results, key_out, key_int):
    outs = [r[key_out] for r in results]
    n = len(outs)
    dev_output = [{w: sum(o[i][w] for o in outs) / n for w in [(0, 0), (1, 1)]} for i in range(len(outs[0]))]
    ints = [r[key_int] for r in results]
    dev_internal = [[{w: sum(x[i][j][w] for x in ints) / n for w in [(0, 0), (1, 1)]} for j in range(len(ints[0][i]))]
                    for i in range(len(ints[0]))]
    return dev_output, dev_internal


def get_avg_output_dev(covariance_results, test_name):
    key = "boost" if test_name.lower().startswith("boost") else "rot"
    return _avg(covariance_results, f"{key}_dev_output", f"{key}_dev_internal")[0]


def get_avg_internal_dev(covariance_results, test_name):
    key = "boost" if test_name.lower().startswith("boost") else "rot"
    return _avg(covariance_results, f"{key}_dev_output", f"{key}_dev_internal")[1]


def display_err(alphas, errs, alpha_name, caption):
    lines = [caption, f"{alpha_name:>14s} {'(0,0)':>14s} {'(1,1)':>14s}"]
    for a, e in zip(alphas, errs):
        lines.append(f"{float(a):14.6g} {e[(0, 0)]:14.4e} {e[(1, 1)]:14.4e}")
    logging.info("\n".join(lines))
    return "\n".join(lines)


def plot_all_dev(dev, save_path):   # plotting is out of scope; kept as a no-op
    return None


def make_dir(path):
    import os
    os.makedirs(path, exist_ok=True)
    return path
