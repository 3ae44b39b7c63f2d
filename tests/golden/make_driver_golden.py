"""Generate tests/golden/driver_golden.npz by RUNNING THE REFERENCE'S OWN DRIVERS (build container).

    python tests/golden/make_driver_golden.py

The unmodified ``CDRNetInferencer`` (inference.py:23-114, configs/mads_3d.yaml) and ``BaseLine``
(baseline.py:22-103, configs/mads_2d.yaml) are constructed exactly as their ``__main__`` does — they
strict-load ``weights/<MODEL.NAME>/best.pth`` / ``latest.pth`` from the working directory — and run
on one synthetic frame (tests/refdrivers.py: driver_case).  Stored: what ``inference`` /
``estimate`` return on the CPU in fp32 (the reference as shipped).  The checkpoints are seeded
random inits (no weights are reachable); this script also asserts that this repo's drop-in modules
create bit-identical parameters under the same seed, so the GPU test can rebuild the checkpoint.
"""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import refdrivers as RD  # noqa: E402
from oracle import refload  # noqa: E402
import fast_3d_human_pose_estimation_b200 as pkg  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "driver_golden.npz")


def main():
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    ref = refload.load()
    inference_mod, baseline_mod = refload.load_drivers()
    case = RD.driver_case()
    G = {}
    cfg3 = refload.load_config("mads_3d.yaml")
    sd = RD.seeded_state_dict(ref.CDRNet, cfg3)
    ours = RD.seeded_state_dict(pkg.CDRNet, cfg3)
    assert list(sd) == list(ours) and all(torch.equal(sd[k], ours[k]) for k in sd), "seeded init differs"
    with tempfile.TemporaryDirectory() as tmp, RD.workdir_with_checkpoint(tmp, cfg3, sd, "best.pth"):
        r = RD.run_cdrnet_driver(inference_mod, cfg3, case)
    for k in ("kp_l", "kp_r", "xyz", "err"):
        G["cdrnet." + k] = r[k]
    print("CDRNetInferencer on", r["device"], "err", r["err"])

    cfg2 = refload.load_config("mads_2d.yaml")
    sd = RD.seeded_state_dict(ref.PoseResNet, cfg2)
    ours = RD.seeded_state_dict(pkg.PoseResNet, cfg2)
    assert list(sd) == list(ours) and all(torch.equal(sd[k], ours[k]) for k in sd), "seeded init differs"
    with tempfile.TemporaryDirectory() as tmp, RD.workdir_with_checkpoint(tmp, cfg2, sd, "latest.pth"):
        r = RD.run_baseline_driver(baseline_mod, cfg2, case)
    for k in ("u8_l", "u8_r", "xyz", "err"):
        G["baseline." + k] = r[k]
    h = r["heat"].reshape(2, r["heat"].shape[1], -1)               # (view, joint, 4096)
    top2 = np.sort(h, axis=2)[:, :, -2:]
    G["baseline.top2_gap"] = top2[:, :, 1] - top2[:, :, 0]         # per view and joint
    G["baseline.heat_absmax"] = np.abs(h).max(2)
    print("BaseLine err", r["err"], "top-2 gaps / |max|", np.sort((G["baseline.top2_gap"] / G["baseline.heat_absmax"]).ravel()))
    np.savez_compressed(OUT, **G)
    print("wrote", OUT, {k: v.shape for k, v in G.items()})


if __name__ == "__main__":
    main()
