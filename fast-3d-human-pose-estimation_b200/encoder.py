"""ResNet backbone — stays on torch/cuDNN (BASELINE.json north_star).

Not part of the hand-written hot path; it only produces the (B,2048,8,8)
latent the head consumes.  The module tree reproduces the reference's
``state_dict`` key names (``conv1``, ``bn1``, ``layer{1..4}.<i>.{conv1,bn1,
conv2,bn2,conv3,bn3,downsample.0,downsample.1}``; reference
models/encoder.py:79-131) so ``best.pth`` / ``latest.pth`` checkpoints load,
and creates parameters in the same order so a seeded random init matches.
"""
from torch import nn

_SPEC = {18: ("basic", (2, 2, 2, 2)), 34: ("basic", (3, 4, 6, 3)),
         50: ("bottleneck", (3, 4, 6, 3)), 101: ("bottleneck", (3, 4, 23, 3)),
         152: ("bottleneck", (3, 8, 36, 3))}


def _bn(c):
    return nn.BatchNorm2d(c, momentum=0.1)


class _Block(nn.Module):
    """Residual block.  ``kind='bottleneck'``: 1x1 -> 3x3(stride) -> 1x1(x4)
    (reference models/encoder.py:38-76).  ``kind='basic'``: two 3x3 convs that
    BOTH carry the stride, as the reference does (models/encoder.py:6-35)."""

    def __init__(self, kind, cin, planes, stride, downsample):
        super().__init__()
        if kind == "bottleneck":
            self.conv1 = nn.Conv2d(cin, planes, 1, bias=False)
            self.bn1 = _bn(planes)
            self.conv2 = nn.Conv2d(planes, planes, 3, stride, 1, bias=False)
            self.bn2 = _bn(planes)
            self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
            self.bn3 = _bn(planes * 4)
        else:
            self.conv1 = nn.Conv2d(cin, planes, 3, stride, 1, bias=False)
            self.bn1 = _bn(planes)
            self.conv2 = nn.Conv2d(planes, planes, 3, stride, 1, bias=False)
            self.bn2 = _bn(planes)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.kind = kind

    def forward(self, x):
        y = self.relu(self.bn1(self.conv1(x)))
        if self.kind == "bottleneck":
            y = self.relu(self.bn2(self.conv2(y)))
            y = self.bn3(self.conv3(y))
        else:
            y = self.bn2(self.conv2(y))
        r = x if self.downsample is None else self.downsample(x)
        return self.relu(y + r)


class ResNet(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        kind, counts = _SPEC[cfg.MODEL.NUM_LAYERS]
        exp = 4 if kind == "bottleneck" else 1
        self.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
        self.bn1 = _bn(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        cin = 64
        for li, (planes, n, stride) in enumerate(
                zip((64, 128, 256, 512), counts, (1, 2, 2, 2)), start=1):
            blocks = []
            for bi in range(n):
                s = stride if bi == 0 else 1
                ds = None
                if bi == 0 and (s != 1 or cin != planes * exp):
                    ds = nn.Sequential(nn.Conv2d(cin, planes * exp, 1, s, bias=False),
                                       _bn(planes * exp))
                blocks.append(_Block(kind, cin, planes, s, ds))
                cin = planes * exp
            setattr(self, f"layer{li}", nn.Sequential(*blocks))
        self.out_channels = cin

    def forward(self, x):
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        return self.layer4(self.layer3(self.layer2(self.layer1(x))))
