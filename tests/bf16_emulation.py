"""Test infrastructure: the oracle with bf16 rounding inserted where the tensor-core path
rounds (weights after BN folding, every activation written to HBM), everything else in fp64.
Lets the GPU tests check the tcgen05 kernels tightly (they then differ only by fp32
accumulation order and rare 1-ulp bf16 ties) instead of only through the loose bf16-vs-fp64
accuracy bounds."""
import torch
import torch.nn.functional as F

from oracle import cdr_oracle as O


def bf(t):
    return t.to(torch.float32).to(torch.bfloat16).to(torch.float64)


def _fold(sd, conv, bn, transposed=False):
    w = sd[conv + ".weight"].double()
    b = sd.get(conv + ".bias")
    b = b.double() if b is not None else torch.zeros(w.shape[1] if transposed else w.shape[0], dtype=torch.float64)
    if bn is None:
        return w, b
    s = sd[bn + ".weight"].double() / torch.sqrt(sd[bn + ".running_var"].double() + O.BN_EPS)
    shape = (1, -1, 1, 1) if transposed else (-1, 1, 1, 1)
    return w * s.reshape(shape), (b - sd[bn + ".running_mean"].double()) * s + sd[bn + ".bias"].double()


def _conv(x, sd, conv, bn, relu=True):
    w, b = _fold(sd, conv, bn)
    y = F.conv2d(x, bf(w), b.float().double())
    return F.relu(y) if relu else y


def decoder_bf16(sd, x, prefix="decoder.", fused_final=False):
    """fused_final: emulate a kernel that fuses final_layer into deconv3's epilogue (deconv3's
    activation never rounded to bf16, fp32 1x1 weights).  Tried on B200 and rejected: 19 x 256 FMAs
    per pixel on the CUDA cores made deconv3's epilogue 3x slower than the MMA main loop."""
    for name in ("deconv1", "deconv2", "deconv3"):
        w, b = _fold(sd, f"{prefix}{name}.0", f"{prefix}{name}.1", transposed=True)
        x = F.relu(F.conv_transpose2d(x, bf(w), b.float().double(), stride=2, padding=1))
        if not (fused_final and name == "deconv3"):
            x = bf(x)
    w, b = _fold(sd, prefix + "final_layer", None)
    return F.conv2d(x, w if fused_final else bf(w), b.float().double())   # heat-maps stay fp32 on the device


def head_bf16(sd, feats, proj_list, pinv_list, taps=None):
    """feats fp32 tensors, proj_list / pinv_list the fp32 matrices the device uses."""
    cs = []
    for x, pinv in zip(feats, pinv_list):
        y = bf(_conv(bf(x), sd, "CF.conv_layer1.0", "CF.conv_layer1.1"))
        cs.append(bf(O.ftl(y, pinv.double())))
    cat = torch.cat(cs, 1)
    f = bf(_conv(cat, sd, "CF.conv_layer2.0", "CF.conv_layer2.1"))
    f = bf(_conv(f, sd, "CF.conv_layer2.3", "CF.conv_layer2.4"))
    hms, kps, fo = [], [], []
    for i, p in enumerate(proj_list):
        g = bf(O.ftl(f, p.double()))
        z = bf(_conv(g, sd, f"CF.out_layer.{i}.0", f"CF.out_layer.{i}.1"))
        fo.append(z)
        h = decoder_bf16(sd, z).float().double()
        hms.append(h)
        kps.append(O.process_heatmap(h) * 4.0)
    projs = torch.stack([p.double() for p in proj_list], 1)
    k = torch.stack(kps, 2)
    xyz = torch.stack([O.dlt(projs, k[:, j]) for j in range(k.shape[1])], 1)
    if taps is not None:
        taps.update(cf_cat=cat, cf_f=f, f_out=fo, heatmaps=hms)
    return kps, xyz


def encoder_bf16(resnet, x):
    """fp64 evaluation of a Bottleneck ``ResNet`` (fast_3d_human_pose_estimation_b200.encoder, module
    tree identical to models/encoder.py:79-131) with bf16 rounding where the tcgen05 encoder rounds:
    BN-folded weights, and every activation written to memory (stem output, conv1/conv2 outputs,
    the downsample branch, the block output after residual + ReLU)."""
    def fold(conv, bn):
        s = bn.weight.double() / torch.sqrt(bn.running_var.double() + bn.eps)
        return conv.weight.double() * s.reshape(-1, 1, 1, 1), bn.bias.double() - bn.running_mean.double() * s

    def cbr(x, conv, bn, relu):
        w, b = fold(conv, bn)
        y = F.conv2d(x, bf(w), b.float().double(), stride=conv.stride, padding=conv.padding)
        return F.relu(y) if relu else y

    r = resnet
    w, b = fold(r.conv1, r.bn1)
    y = bf(F.relu(F.conv2d(bf(x), bf(w), b.float().double(), stride=r.conv1.stride, padding=r.conv1.padding)))
    y = F.max_pool2d(y, 3, 2, 1)
    for li in range(1, 5):
        for blk in getattr(r, f"layer{li}"):
            t = bf(cbr(y, blk.conv1, blk.bn1, True))
            t = bf(cbr(t, blk.conv2, blk.bn2, True))
            res = y if blk.downsample is None else bf(cbr(y, blk.downsample[0], blk.downsample[1], False))
            y = bf(F.relu(cbr(t, blk.conv3, blk.bn3, False) + res))
    return y
