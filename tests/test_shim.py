"""CPU: the Python mirror of the reference interface — state_dict contract, construction order,
error behaviour.  The product path must refuse to run without CUDA instead of falling back."""
import numpy as np
import pytest
import torch

from fast_3d_human_pose_estimation_b200 import synth

HEAD_KEYS = {
    "CF.conv_layer1.0.weight": (300, 2048, 1, 1), "CF.conv_layer1.0.bias": (300,),
    "CF.conv_layer1.1.running_var": (300,),
    "CF.conv_layer2.0.weight": (400, 800, 1, 1), "CF.conv_layer2.3.weight": (400, 400, 1, 1),
    "CF.conv_layer2.4.running_mean": (400,),
    "CF.out_layer.0.0.weight": (2048, 300, 1, 1), "CF.out_layer.1.1.bias": (2048,),
    "decoder.deconv1.0.weight": (2048, 256, 4, 4), "decoder.deconv2.0.weight": (256, 256, 4, 4),
    "decoder.deconv3.1.running_var": (256,), "decoder.final_layer.weight": (19, 256, 1, 1),
    "decoder.final_layer.bias": (19,),
}


def test_state_dict_contract(pkg):
    m = pkg.CDRNet(synth.make_cfg(18, 19))
    sd = m.state_dict()
    for k, shape in HEAD_KEYS.items():
        assert tuple(sd[k].shape) == shape, k
    assert "encoder.layer4.1.conv2.weight" in sd and "encoder.conv1.weight" in sd
    assert not any(k.startswith("_") for k in sd)
    head = synth.make_head_state_dict()
    missing, unexpected = m.load_state_dict(head, strict=False)
    assert not unexpected and all(k.startswith("encoder.") for k in missing)
    p = pkg.PoseResNet(synth.make_cfg(18, 16))
    assert tuple(p.state_dict()["decoder.final_layer.weight"].shape) == (16, 256, 1, 1)


def test_no_cpu_fallback(pkg):
    m = pkg.CDRNet(synth.make_cfg(18, 19)).eval()
    feats = synth.make_features(1)
    cams = synth.make_cameras(1)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m.head(feats, [torch.from_numpy(cams["P_l"]), torch.from_numpy(cams["P_r"])])
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="CUDA"):
            pkg.calc_mpjpe([np.zeros((19, 2)), np.zeros((19, 2))], np.zeros((19, 3)), np.zeros((19, 3)),
                           np.zeros((19, 2)), np.zeros((19, 2)))
        with pytest.raises(RuntimeError, match="CUDA"):
            pkg.get_max_preds(np.zeros((1, 19, 64, 64), np.float32))
        with pytest.raises(RuntimeError, match="CUDA"):
            pkg.triangulation(np.eye(4), np.eye(4), np.zeros((19, 2), np.uint8), np.zeros((19, 2), np.uint8))


def test_training_mode_rejected(pkg):
    m = pkg.CDRNet(synth.make_cfg(18, 19))
    assert m.training
    with pytest.raises(RuntimeError, match="inference-only"):
        m.head([torch.zeros(1, 2048, 8, 8)] * 2, [torch.zeros(1, 3, 4)] * 2)


def test_reference_argument_errors(pkg):
    with pytest.raises(NotImplementedError):
        pkg.CDRNet(synth.make_cfg(18, 19), n_views=3)
    with pytest.raises(NotImplementedError):
        pkg.CDRNet(synth.make_cfg(18, 19), fusion_in_dim=512)
    with pytest.raises(ValueError, match="4/3"):          # the reference's ftl cannot reshape 2*300 channels into 4 blocks of 3
        pkg.CDRNet(synth.make_cfg(18, 19), fusion_hid_ch1=300, fusion_hid_ch2=300)
    with pytest.raises(NotImplementedError):
        pkg.CDRNet(synth.make_cfg(18, 19), fusion_hid_ch1=30, fusion_hid_ch2=40)      # not a multiple of 12
    m = pkg.CDRNet(synth.make_cfg(18, 19), fusion_hid_ch1=192, fusion_hid_ch2=256)
    assert m.CF.conv_layer1[0].out_channels == 192 and m.CF.conv_layer2[0].in_channels == 512
    with pytest.raises(ValueError):
        pkg.CDRNet(synth.make_cfg(18, 19), precision="bf16", encoder_precision="fp32")   # fp16-plane latents need the fp32 head
    with pytest.raises(ValueError):
        pkg.CDRNet(synth.make_cfg(18, 19), precision="fp8")
    with pytest.raises(ValueError, match="does not exist"):
        pkg.CDRNet(synth.make_cfg(18, 19)).init_weights("/nonexistent.pth")
    with pytest.raises(AssertionError):
        if torch.cuda.is_available():
            pkg.get_max_preds(np.zeros((19, 64, 64), np.float32))
        else:
            raise AssertionError


def test_synth_is_deterministic():
    a, b = synth.make_features(2, seed=1), synth.make_features(2, seed=1)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    assert np.array_equal(synth.make_cameras(3, seed=2)["P_l"], synth.make_cameras(3, seed=2)["P_l"])
    s1, s2 = synth.make_head_state_dict(seed=0), synth.make_head_state_dict(seed=0)
    assert all(torch.equal(s1[k], s2[k]) for k in s1)


def test_autograd_ops_refuse_cpu_tensors(pkg):
    """SURVEY §8f rank 3 slice: the differentiable ops run on the library only — CPU tensors are an error, never a
    silent torch fallback."""
    with pytest.raises(TypeError, match="no CPU fallback"):
        pkg.soft_argmax_2d(torch.zeros(1, 2, 64, 64), 4.0)
    with pytest.raises(TypeError, match="no CPU fallback"):
        pkg.dlt(torch.zeros(1, 3, 4), torch.zeros(1, 3, 4), torch.zeros(1, 2, 2), torch.zeros(1, 2, 2))
    with pytest.raises(TypeError, match="no CPU fallback"):
        pkg.ftl(torch.zeros(1, 300, 8, 8), torch.zeros(1, 4, 3))


def test_numa_binding_is_harmless(pkg):
    import os
    from fast_3d_human_pose_estimation_b200 import dist
    before = os.sched_getaffinity(0)
    cpus = dist.bind_to_gpu_numa(0)            # None without NVML / a GPU; a subset of the allowed CPUs otherwise
    assert cpus is None or set(cpus) <= before
    os.sched_setaffinity(0, before)


def test_checkpoint_contract_matches_reference_exactly(pkg, tmp_path):
    """SURVEY §8f rank 4 (checkpoint I/O): `inference.py:31-33` loads `best.pth` with a strict
    `load_state_dict`, `train_cdr.py:223-232` saves `model.state_dict()`.  The drop-in modules expose the
    reference's keys, shapes and dtypes in the reference's order (fixture generated from the reference by
    tests/golden/make_state_dict_keys.py), and a saved checkpoint round-trips."""
    import json
    import os
    want = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_state_dict_keys.json")))

    def describe(m):
        return [[k, list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in m.state_dict().items()]
    assert describe(pkg.CDRNet(synth.make_cfg(50, 19))) == want["CDRNet50"]
    m = pkg.CDRNet(synth.make_cfg(101, 19))
    assert describe(m) == want["CDRNet101"]
    assert describe(pkg.PoseResNet(synth.make_cfg(101, 16))) == want["PoseResNet101_j16"]
    # a reference-format checkpoint (plain state_dict, as train_cdr.py writes it) loads strictly, also into the
    # variants with extra constructor keywords
    path = tmp_path / "best.pth"
    torch.save(m.state_dict(), path)
    for kw in ({}, {"precision": "bf16"}, {"encoder_precision": "bf16"}, {"trainable": True}):
        m2 = pkg.CDRNet(synth.make_cfg(101, 19), **kw)
        m2.load_state_dict(torch.load(path), strict=True)
        assert torch.equal(m2.state_dict()["CF.out_layer.1.0.weight"], m.state_dict()["CF.out_layer.1.0.weight"])
