"""Golden anomaly scores from the UNMODIFIED reference (utils/jet_analysis/anomaly_detection.py:251-419), build container only.

    python tests/golden/make_scores_golden.py     # writes tests/golden/anomaly_scores.pt

matplotlib / jetnet / energyflow / awkward / coffea (absent here, unused by the functions called) come from baseline/stubs."""
import os
import sys

REF = os.environ.get("LGAE_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path = [p for p in sys.path if os.path.abspath(p or ".") != ROOT]
sys.path.insert(0, os.path.join(ROOT, "baseline", "stubs"))
sys.path.insert(0, REF)

import torch  # noqa: E402

import lgn.models  # noqa: E402,F401
from utils.jet_analysis import anomaly_detection as ad  # noqa: E402

assert ad.__file__.startswith(REF)
g = torch.Generator().manual_seed(23)
B, N = 5, 30
target = torch.rand(B, N, 4, generator=g, dtype=torch.float64)
target[..., 0] += 1.0
recons = target + 0.05 * torch.randn(B, N, 4, generator=g, dtype=torch.float64)
perm = torch.stack([torch.randperm(N, generator=g) for _ in range(B)])
recons = torch.gather(recons, 1, perm[:, :, None].expand(B, N, 4))
jet = lambda p: p.sum(-2)
out = {
    "recons": recons, "target": target,
    "chamfer_cartesian": ad.chamfer(recons, target).mean(-1),
    "mse_cartesian": ad.mse(recons, target).mean(-1),
    "chamfer_lorentz": ad.chamfer_lorentz(recons, target).mean(-1),
    "mse_lorentz": ad.mse_lorentz(recons, target).mean(-1),
    "jet_cartesian": ad.mse(jet(recons), jet(target)),
    "jet_lorentz": ad.mse_lorentz(jet(recons), jet(target)),
}
torch.save(out, os.path.join(HERE, "anomaly_scores.pt"))
print({k: v.shape for k, v in out.items()})
