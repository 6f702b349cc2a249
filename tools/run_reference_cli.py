#!/usr/bin/env python
"""Run the reference's own, UNCHANGED scripts (main.py -> test.py -> covariance_test.py) against this repository's drop-in
`lgn` package (the north-star's boundary: "drops into main.py/test.py unchanged"), or -- `--impl reference` -- against the
reference's own `lgn/` for a side-by-side log.

    python tools/setup_reference.py                      # build container: copies /root/reference to baseline/_ref (git-ignored)
    python tools/run_reference_cli.py --out gpurun_out/refcli --device cuda          # B200 box: the shim
    python tools/run_reference_cli.py --out gpurun_out/refcli_ref --impl reference --device cpu

The scripts run with cwd = baseline/_ref (so `utils.*` resolves to the reference's), PYTHONPATH = <repo>:<repo>/baseline/stubs for
the shim (a regular package `lgn/__init__.py` shadows the reference's namespace package) or just the stubs for
the reference.  Data: synthetic jets written in the reference's .pt format {'p4','labels','Nobj'} (utils/data/preprocess.py:73-90).
Nothing is patched: the stubs only stand in for matplotlib / jetnet / energyflow / awkward / coffea, which this image lacks."""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
STUBS = os.path.join(ROOT, "baseline", "stubs")


def write_data(path, jets, n, seed):
    import torch
    sys.path.insert(0, ROOT)
    from oracle.lgae_oracle import synthetic_jets   # the same generator as the parity tests (test infrastructure)
    d = synthetic_jets(jets, n, seed=seed, mass_scale=1e-6, pad=True)
    torch.save({"p4": d["p4"], "labels": d["labels"], "Nobj": d["Nobj"]}, path)


def run(cmd, env, cwd, log):
    t0 = time.time()
    with open(log, "w") as f:
        p = subprocess.run(cmd, env=env, cwd=cwd, stdout=f, stderr=subprocess.STDOUT)
    return p.returncode, time.time() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "refcli"))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--jets", type=int, default=2048)
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--test-jets", type=int, default=64)
    ap.add_argument("--skip", default="", help="comma list of stages to skip: main,test,cov")
    args = ap.parse_args()
    if not os.path.isdir(os.path.join(REF, "lgn")):
        raise SystemExit("baseline/_ref is missing: run tools/setup_reference.py in the build container first")
    out = os.path.abspath(args.out)
    os.makedirs(out, exist_ok=True)
    data = os.path.join(out, "jets_train.pt")
    test_data = os.path.join(out, "jets_test.pt")
    write_data(data, args.jets, 30, seed=1)
    write_data(test_data, args.test_jets, 30, seed=2)
    env = dict(os.environ)
    env["PYTHONPATH"] = (ROOT + os.pathsep if args.impl == "b200" else "") + STUBS
    env["MPLBACKEND"] = "Agg"
    save = os.path.join(out, "exp")
    model = ["-j", "QCD", "--maxdim", "2", "--tau-latent-vectors", "8", "--tau-latent-scalars", "1", "--map-to-latent", "min&max",
             "--mlp-width", "6", "--mlp-depth", "6", "--encoder-num-channels", "3", "3", "4", "4", "--decoder-num-channels", "4", "4", "3", "3",
             "--device", args.device]
    test_dev = ["--test-device", args.device]
    skip = set(args.skip.split(","))
    summary = {"impl": args.impl, "device": args.device}
    py = [sys.executable, "-u"]
    if "main" not in skip:
        cmd = py + ["main.py", "--data-paths", data, "--test-data-paths", test_data, "-e", str(args.epochs), "-bs", str(args.batch),
                    "--train-fraction", "0.75", "--lr", "0.0005", "--loss-choice", "chamfer", "--get-real-method", "sum", "--l1-lambda", "1e-8",
                    "--l2-lambda", "0", "--patience", "1000", "--plot-freq", "1000", "--save-freq", "1", "--plot-start-epoch", "1000",
                    "--equivariance-test", "--num-test-batch", "1", "--test-batch-size", "16", "--save-dir", save, "--seed", "0"] + model + test_dev
        rc, dt = run(cmd, env, REF, os.path.join(out, "main.log"))
        summary["main"] = {"rc": rc, "seconds": dt}
    # the folder main.py created
    exp = None
    if os.path.isdir(save):
        subs = sorted(os.path.join(save, d) for d in os.listdir(save))
        exp = subs[-1] if subs else None
    summary["model_path"] = exp
    if "test" not in skip and exp:
        cmd = py + ["test.py", "--test-data-paths", test_data, "--model-path", exp, "--test-batch-size", "32", "--get-real-method", "sum",
                    "--loss-choice", "chamfer", "--plot-freq", "1000"] + model + test_dev
        rc, dt = run(cmd, env, REF, os.path.join(out, "test.log"))
        summary["test"] = {"rc": rc, "seconds": dt}
    if "cov" not in skip and exp:
        cmd = py + ["covariance_test.py", "--test-data-paths", test_data, "--model-path", exp, "--test-batch-size", "16", "--num-test-batch", "1"] + model
        rc, dt = run(cmd, env, REF, os.path.join(out, "covariance_test.log"))
        summary["covariance_test"] = {"rc": rc, "seconds": dt}
    with open(os.path.join(out, "summary.json"), "w") as f:
        json.dump(summary, f, indent=1)
    print(json.dumps(summary))
    for k in ("main", "test", "covariance_test"):
        if k in summary and summary[k]["rc"] != 0:
            raise SystemExit(f"{k} failed: see {out}/{k}.log")


if __name__ == "__main__":
    main()
