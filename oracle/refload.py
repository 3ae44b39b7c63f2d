"""Import the UNMODIFIED reference from /root/reference (build container only).

TEST INFRASTRUCTURE.  /root/reference does not exist on the GPU box, so only
``tests/golden/make_golden.py``, ``oracle/validate_against_reference.py`` and
tests that skip when it is absent use this.  ``matplotlib`` (imported by the
reference's tools/utils.py:3-4,10 for plotting only) is not installed here, so a
stub module is registered before the import; nothing on the hot path touches it.
"""
from __future__ import annotations

import os
import sys
import types

REF_ROOT = os.environ.get("CDR_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "models", "cdrnet.py"))


def load():
    """Returns a namespace with the reference's CDRNet, PoseResNet, calc_mpjpe,
    get_max_preds, triangulation, get_projection_matrix, project_3d_to_2d."""
    if not available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    if "matplotlib" not in sys.modules:
        m = types.ModuleType("matplotlib")
        m.use = lambda *a, **k: None
        p = types.ModuleType("matplotlib.pyplot")
        m.pyplot = p
        sys.modules["matplotlib"] = m
        sys.modules["matplotlib.pyplot"] = p
    # the reference's top-level packages are called `models` / `tools`; import them under
    # a clean path so they cannot shadow anything of ours
    saved = list(sys.path)
    sys.path.insert(0, REF_ROOT)
    try:
        from models.cdrnet import CDRNet
        from models.poseresnet import PoseResNet
        from models.metrics import calc_mpjpe
        from tools.utils import get_max_preds
        from tools.common import triangulation, get_projection_matrix, project_3d_to_2d
    finally:
        sys.path[:] = saved
    return types.SimpleNamespace(
        CDRNet=CDRNet, PoseResNet=PoseResNet, calc_mpjpe=calc_mpjpe, get_max_preds=get_max_preds,
        triangulation=triangulation, get_projection_matrix=get_projection_matrix,
        project_3d_to_2d=project_3d_to_2d)


def feature_stub(feats):
    """Stands in for ``model.encoder`` so the reference's own ``CDRNet.forward`` runs its
    post-encoder path (models/cdrnet.py:236-268) on given latents: returns the queued
    features in call order (left, right) while the image tensors still set img_size."""
    import torch

    class FeatureStub(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.feats = list(feats)
            self.i = 0

        def forward(self, x):
            f = self.feats[self.i % len(self.feats)]
            self.i += 1
            return f

    return FeatureStub()
