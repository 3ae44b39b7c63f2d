"""Robustness of the shim around the kernels: per-stream workspaces, graph-owned scratch,
packed-weight lifetimes under captured graphs, several devices in one process."""
import copy

import numpy as np
import pytest
import torch

from fast_3d_human_pose_estimation_b200 import synth, workspace

pytestmark = pytest.mark.gpu


def _model(pkg, sd, precision="fp32", device="cuda"):
    m = pkg.CDRNet(synth.make_cfg(18, 19), precision=precision)
    m.load_state_dict(sd, strict=False)
    return m.to(device).eval()


def _inputs(b, seed, device="cuda"):
    feats = [f.to(device) for f in synth.make_features(b, seed=seed)]
    cams = synth.make_cameras(b, seed=seed + 1)
    return feats, [torch.from_numpy(cams["P_l"]).to(device), torch.from_numpy(cams["P_r"]).to(device)]


def test_two_models_on_two_streams_do_not_share_scratch(cuda_pkg):
    """Two models running concurrently on two streams: each stream has its own workspace, so the results
    equal the serial ones bit for bit (with ONE process-global buffer they corrupt each other)."""
    sd_a = synth.make_head_state_dict(seed=0, calibrated=True)
    sd_b = synth.make_head_state_dict(seed=5, calibrated=True)
    ma, mb = _model(cuda_pkg, sd_a), _model(cuda_pkg, sd_b, precision="bf16")
    fa, Pa = _inputs(6, 1)
    fb, Pb = _inputs(4, 7)
    (wa, _), xa = ma.head(fa, Pa)
    (wb, _), xb = mb.head(fb, Pb)
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs = []
    for _ in range(6):
        with torch.cuda.stream(s1):
            ra = ma.head(fa, Pa)
        with torch.cuda.stream(s2):
            rb = mb.head(fb, Pb)
        outs.append((ra, rb))
    torch.cuda.synchronize()
    for ((ka, _), x1), ((kb, _), x2) in outs:
        assert torch.equal(ka, wa) and torch.equal(x1, xa) and torch.equal(kb, wb) and torch.equal(x2, xb)
    assert len(workspace._DEFAULT) >= 3                  # default stream + the two side streams


def test_graph_owns_its_workspace_and_survives_larger_eager_calls(cuda_pkg):
    sd = synth.make_head_state_dict(seed=0, calibrated=True)
    m = _model(cuda_pkg, sd)
    b = 2
    feats_h = [f.pin_memory() for f in synth.make_features(b, seed=1)]
    cams = synth.make_cameras(b, seed=2)
    P_h = [torch.from_numpy(cams["P_l"]).pin_memory(), torch.from_numpy(cams["P_r"]).pin_memory()]
    hg = cuda_pkg.HeadGraph(m, feats_h, P_h)
    kp0, xyz0, _ = hg.replay()
    kp0, xyz0 = [k.clone() for k in kp0], xyz0.clone()
    big_f, big_P = _inputs(9, 3)                         # grows the eager (per-stream) workspace, not the graph's
    m.head(big_f, big_P)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):                        # eager work on another stream while the graph replays
        for _ in range(3):
            m.head(big_f, big_P)
    kp1, xyz1, _ = hg.replay()
    torch.cuda.synchronize()
    assert torch.equal(kp1[0], kp0[0]) and torch.equal(xyz1, xyz0)
    assert hg._ws.nbytes() > 0 and all(hg._ws is not w for w in workspace._DEFAULT.values())


def test_graph_refuses_stale_weights_and_keeps_them_alive(cuda_pkg):
    """After a capture: a parameter update + eager forward re-packs the weights.  The graph's own reference keeps
    the old pool alive (no use-after-free) and its next replay raises instead of silently using stale weights."""
    sd = synth.make_head_state_dict(seed=0, calibrated=True)
    m = _model(cuda_pkg, sd)
    pipe = cuda_pkg.HeadPipeline(m, 2)
    feats_h = [f.pin_memory() for f in synth.make_features(2, seed=1)]
    cams = synth.make_cameras(2, seed=2)
    P_h = [torch.from_numpy(cams["P_l"]).pin_memory(), torch.from_numpy(cams["P_r"]).pin_memory()]
    pipe.submit(feats_h, P_h)
    _, kp, xyz, _ = pipe.collect()
    old_box = m._packed._box
    with torch.no_grad():
        m.decoder.final_layer.weight.mul_(1.5)
    f, P = _inputs(2, 1)
    m.head(f, P)                                         # re-pack: a new handle for the module ...
    torch.cuda.synchronize()
    assert m._packed._box is not old_box and old_box.handle is not None and old_box.refs == 1   # ... the graph still holds the old one
    with pytest.raises(RuntimeError, match="parameters changed"):
        pipe.submit(feats_h, P_h)
    pipe.close()
    assert old_box.handle is None                        # last reference gone -> destroyed
    pipe2 = cuda_pkg.HeadPipeline(m, 2)                  # a fresh pipeline sees the new weights
    pipe2.submit(feats_h, P_h)
    _, kp2, xyz2, _ = pipe2.collect()
    assert not torch.equal(kp2[0], kp[0])


def test_repack_inside_capture_is_refused(cuda_pkg):
    sd = synth.make_head_state_dict(seed=0, calibrated=True)
    m = _model(cuda_pkg, sd)
    f, P = _inputs(1, 1)
    g = torch.cuda.CUDAGraph()
    with pytest.raises(RuntimeError, match="captur"):
        with torch.cuda.graph(g):
            m.head(f, P)                                 # never packed: packing would cudaMalloc + synchronise
    torch.cuda.synchronize()
    m.head(f, P)                                         # and the model is still usable afterwards
    torch.cuda.synchronize()


def test_deepcopy_and_nondefault_bn_eps(cuda_pkg):
    torch.manual_seed(0)
    m = cuda_pkg.CDRNet(synth.make_cfg(50, 19), encoder_precision="bf16")
    m.load_state_dict(synth.make_head_state_dict(seed=0, calibrated=True), strict=False)
    m = m.cuda().eval()
    xs = [x.cuda() for x in synth.make_images(1, seed=1)]
    _, P = _inputs(1, 1)
    (k1, _), x1 = m(xs, P)
    m2 = copy.deepcopy(m)                                # packed handles are not copied or shared
    (k2, _), x2 = m2(xs, P)
    del m
    torch.cuda.synchronize()
    (k3, _), x3 = m2(xs, P)
    assert torch.equal(k1, k2) and torch.equal(x1, x3)
    m2.decoder.deconv1[1].eps = 1e-3
    with torch.no_grad():
        m2.decoder.deconv1[1].weight.add_(0.0)           # bump the version -> re-pack
    with pytest.raises(ValueError, match="eps"):
        m2.head([torch.zeros(1, 2048, 8, 8, device="cuda")] * 2, P)
    mean = torch.tensor([0.5, 0.5, 0.5])                 # tensor mean/std (ADVICE: `mean or DEFAULT` raised)
    frames = torch.randint(0, 256, (2, 256, 256, 3), dtype=torch.uint8, device="cuda")
    m3 = copy.deepcopy(m2)
    m3.decoder.deconv1[1].eps = 1e-5
    out = m3.forward_frames(frames, P, mean=mean, std=np.array([0.25, 0.25, 0.25]))
    assert torch.isfinite(out[1]).all()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_second_device_in_one_process(cuda_pkg):
    """Kernel attributes (> 48 KB dynamic shared memory) and the SM count are per device: a model on cuda:1
    after one on cuda:0 must launch and agree bit for bit."""
    sd = synth.make_head_state_dict(seed=0, calibrated=True)
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        m = _model(cuda_pkg, sd, device=dev)
        f, P = _inputs(3, 1, device=dev)
        (kl, _), xyz = m.head(f, P)
        hm = cuda_pkg.PoseDecoder(synth.make_cfg(18, 19)).to(dev).eval()(f[0])
        pts = cuda_pkg.baseline_keypoints(hm)
        torch.cuda.synchronize(dev)
        outs.append((kl.cpu(), xyz.cpu(), pts.cpu()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


def test_pipeline_rows_latents_and_gather_buffer(cuda_pkg):
    """HeadPipeline fed bf16 latent rows (half the H2D bytes) == the eager head on the same rows; with gather_total the 3D
    joints and MPJPE sums land in this rank's slot of a dist.GatherBuffer (world 1 here: the collective is a no-op, the
    in-place plumbing and the host-side unpack are what is checked; NCCL at N > 1: bench.py's sharded_check)."""
    b = 3
    sd = synth.make_head_state_dict(seed=0, calibrated=True)
    m = _model(cuda_pkg, sd, precision="bf16")
    feats = synth.make_features(b, seed=4)
    cams = synth.make_cameras(b, seed=5)
    gt = synth.make_gt(cams, seed=6)
    gtd = {k: torch.from_numpy(gt[k]).cuda() for k in ("gt3d", "gt2d_l", "gt2d_r", "vis")}
    P_h = [torch.from_numpy(cams["P_l"]).pin_memory(), torch.from_numpy(cams["P_r"]).pin_memory()]
    rows_h = torch.cat([f.permute(0, 2, 3, 1).reshape(-1, 2048) for f in feats], 0).to(torch.bfloat16).contiguous().pin_memory()
    (kl, kr), xyz = m.head(None, [p.cuda() for p in P_h], feat_rows=rows_h.cuda())
    sums = cuda_pkg.mpjpe_sums([kl, kr], xyz, gtd["gt3d"], gtd["gt2d_l"], gtd["gt2d_r"], gtd["vis"])
    torch.cuda.synchronize()
    pipe = cuda_pkg.HeadPipeline(m, b, gt=gtd, latents="rows_bf16")
    pipe.submit(rows_h, P_h)
    _, kp_h, xyz_h, sums_h = pipe.collect()
    assert torch.equal(kp_h[0], kl.cpu()) and torch.equal(xyz_h, xyz.cpu()) and torch.equal(sums_h, sums.cpu())
    pipe2 = cuda_pkg.HeadPipeline(m, b, gt=gtd, latents="rows_bf16", gather_total=b)
    for _ in range(2):
        pipe2.submit(rows_h, P_h)
        _, kp_h, gbuf, _ = pipe2.collect()
        gx, gs = gbuf.unpack_host()
        assert torch.equal(gx, xyz.cpu()) and torch.equal(gs, sums.cpu()) and torch.equal(kp_h[1], kr.cpu())
