"""Golden equivariance tables from the UNMODIFIED reference harness (lgn/models/autotest/lgn_tests.py), run in the
build container only:   python tests/golden/make_equivariance_golden.py   -> tests/golden/equivariance_cfg1.pt

The fixture holds the jets, both state_dicts and the reference's own boost / rotation deviation tables
(mean-based metric of autotest/utils.py:34-42) for cfg-1 at random init, so that the GPU test can require the
B200 path to be "no worse than the reference" on identical inputs and weights."""
import os
import sys
import types

REF = os.environ.get("LGAE_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path = [p for p in sys.path if os.path.abspath(p or ".") != os.path.abspath(os.path.join(HERE, "..", ".."))]
sys.path.insert(0, REF)


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Stub(self.__name__ + "." + name)

    def __call__(self, *a, **k):
        return None


for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.ticker", "matplotlib.cm", "jetnet", "jetnet.losses"):
    sys.modules.setdefault(name, _Stub(name))

import torch  # noqa: E402

import lgn.models  # noqa: E402,F401
from lgn.models.autotest.lgn_tests import covariance_test  # noqa: E402

sys.path.insert(0, HERE)
from make_golden import build, synthetic_jets  # noqa: E402

assert lgn.models.__file__.startswith(REF)


def main():
    cfg = dict(seed=0, batch=8, n=30, maxdim=2, enc_channels=[3, 3, 4, 4], dec_channels=[4, 4, 3, 3], tau_s=1, tau_v=8,
               map_to_latent="min&max", mass_scale=1e-6, pad=False)
    enc, dec = build(cfg)
    enc.eval(); dec.eval()
    jets = synthetic_jets(cfg["batch"], cfg["n"], seed=11, mass_scale=cfg["mass_scale"])
    data = {"p4": jets["p4"].clone()}
    boost = covariance_test(enc, dec, {"p4": data["p4"].clone()}, "boost", axis="z", unit="TeV")
    rot = covariance_test(enc, dec, {"p4": data["p4"].clone()}, "rotation", axis="z", unit="TeV")
    out = {
        "cfg": cfg, "p4": data["p4"],
        "enc_state": {k: v.detach().clone() for k, v in enc.state_dict().items()},
        "dec_state": {k: v.detach().clone() for k, v in dec.state_dict().items()},
        "gammas": [float(g) for g in boost["gammas"]], "thetas": [float(t) for t in rot["thetas"]],
        "boost_dev_output": [{str(k): float(v) for k, v in d.items()} for d in boost["boost_dev_output"]],
        "rot_dev_output": [{str(k): float(v) for k, v in d.items()} for d in rot["rot_dev_output"]],
        "boost_dev_internal": [[{str(k): float(v) for k, v in d.items()} for d in lvl] for lvl in boost["boost_dev_internal"]],
        "rot_dev_internal": [[{str(k): float(v) for k, v in d.items()} for d in lvl] for lvl in rot["rot_dev_internal"]],
        "torch_version": torch.__version__,
    }
    torch.save(out, os.path.join(HERE, "equivariance_cfg1.pt"))
    for g, d in list(zip(out["gammas"], out["boost_dev_output"]))[::5]:
        print(f"gamma {g:10.3f}  {d}")
    for t, d in list(zip(out["thetas"], out["rot_dev_output"]))[::5]:
        print(f"theta {t:10.3f}  {d}")


if __name__ == "__main__":
    main()
