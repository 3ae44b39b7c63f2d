"""Generate tests/golden/reference_state_dict_keys.json by IMPORTING THE REFERENCE (build container only).

    python tests/golden/make_state_dict_keys.py

The persistence contract of SURVEY Appendix B / §8f rank 4: `inference.py:31-33` does
`model.load_state_dict(torch.load("weights/<NAME>/best.pth"))` (strict), so the drop-in CDRNet /
PoseResNet must expose exactly the reference's keys, shapes and dtypes, in the same order.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import refload  # noqa: E402
from fast_3d_human_pose_estimation_b200 import synth  # noqa: E402


def describe(m):
    return [[k, list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in m.state_dict().items()]


def main():
    ref = refload.load()
    out = {}
    for layers in (50, 101):
        out[f"CDRNet{layers}"] = describe(ref.CDRNet(synth.make_cfg(layers, 19)))
    out["PoseResNet101_j16"] = describe(ref.PoseResNet(synth.make_cfg(101, 16)))
    path = os.path.join(ROOT, "tests", "golden", "reference_state_dict_keys.json")
    json.dump(out, open(path, "w"))
    print("wrote", path, {k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
