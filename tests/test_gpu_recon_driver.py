"""The per-batch driver of fit_recon (vistracker_b200/recon_driver.py): chunked filter == one filter call, mini-batched neural
reconstruction, and one short end-to-end pass through optimize_smpl + the three object phases."""
import numpy as np
import pytest
import torch

from recon_problem import make_problem

pytestmark = pytest.mark.gpu


def _setup(B):
    from scipy.spatial import ConvexHull
    from vistracker_b200 import CHORETriplaneVisibility, default_options, resolve_dims
    from vistracker_b200.generator import GeneratorTriplaneVis
    from vistracker_b200.recon_fit import Priors, ReconFitterTriVisFull, SMPLParams
    from vistracker_b200.render import SilLossROI
    from vistracker_b200.smpl import LandmarkRegressor, SMPL_Layer
    from vistracker_b200.synth import synthetic_frames, synthetic_state_dict
    dev = torch.device("cuda", 0)
    pb = make_problem(seed=31)
    net = CHORETriplaneVisibility(default_options(), device=dev).eval()
    net.load_state_dict(synthetic_state_dict(resolve_dims(default_options()), seed=0))
    net.defer_checks = True
    layer = SMPL_Layer.from_buffers(pb["model"], pb["model"]["parents"], dev)
    reg = pb["reg"]
    body25 = LandmarkRegressor(np.stack([reg[0], reg[1]]), reg[2], reg[3], dev)
    fitter = ReconFitterTriVisFull(net, Priors(pb["assets"], dev), pb["labels"])
    images, _, crop, body = synthetic_frames(B, size=512, seed=3, n_points=4, jitter=True)
    rep = lambda t: t.repeat((B + 3) // 4, *([1] * (t.dim() - 1)))[:B]
    rng = np.random.Generator(np.random.PCG64(0))
    p = rng.standard_normal((400, 3)); p /= np.linalg.norm(p, axis=1, keepdims=True); p *= np.array([0.3, 0.25, 0.2])
    faces = ConvexHull(p[:200]).simplices
    K = SilLossROI.compute_K_roi((424.0, 168.0, 1200.0, 1200.0))[None].repeat(B, 1, 1)
    ref = torch.zeros(B, 256, 256); ref[:, 80:176, 96:160] = 1
    sil = SilLossROI(torch.ones(B, 256, 256), ref, K, p[:200].astype(np.float32), faces, rend_size=256, device=dev)
    smpl_init = lambda human_t: SMPLParams(layer, body25, rep(pb["pose"]), rep(pb["betas"]), human_t.to(dev))
    data = {"images": images, "crop_center": crop, "body_center": body}
    gen = GeneratorTriplaneVis(net, threshold=2.0, filter_val=10.0)                 # random-init UDF: accept every in-front point as surface
    return dev, net, fitter, gen, data, smpl_init, rep(pb["body_kpts"]), torch.from_numpy(p.astype(np.float32)), sil


def test_chunked_filter_equals_one_call():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from vistracker_b200.recon_driver import filter_batch
    dev, net, *_rest = _setup(5)
    images = _rest[2]["images"]
    net.filter(images.to(dev))
    whole = [m.clone() for m in net._maps]
    filter_batch(net, images, chunk=2)                                             # 2 + 2 + 1 frames
    for a, b in zip(whole, net._maps):                                             # GroupNorm statistics are summed with atomics: equal to rounding
        assert a.shape == b.shape and float((a - b).abs().max()) <= 1e-5 * float(a.abs().max())


def test_maps_kept_from_the_generator_equal_a_second_filter():
    """fit_recon_batch keeps the feature maps of the generator's per-mini-batch filter calls instead of filtering the whole batch again
    (recon_fit_triplane.py:57-60): both give the same maps (GroupNorm statistics are summed with atomics: equal to rounding)."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from vistracker_b200.recon_driver import filter_batch, generate_all
    dev, net, fitter, gen, data, *_rest = _setup(5)
    torch.manual_seed(0)
    generate_all(gen, data, mini_batch_size=2, keep_maps=True)
    kept = [m.clone() for m in net._maps]
    filter_batch(net, data["images"], chunk=2)
    for a, b in zip(kept, net._maps):
        assert a.shape == b.shape and float((a - b).abs().max()) <= 1e-5 * float(b.abs().max())


def test_fit_recon_batch_runs_end_to_end():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from vistracker_b200.recon_driver import fit_recon_batch, scale_body_kpts
    B = 5
    dev, net, fitter, gen, data, smpl_init, kpts, obj_points, sil = _setup(B)
    seen = []
    torch.manual_seed(0)
    only = fit_recon_batch(fitter, gen, data, smpl_init, kpts, obj_points, neural_only=True, mini_batch_size=2,
                           on_mini_batch=lambda s, e, pc: seen.append((s, e, pc["object"]["pca_axis"].shape[0])))
    assert seen == [(0, 2, 2), (2, 4, 2), (4, 5, 1)] and set(only) == {"pc_generated"}
    pc = only["pc_generated"]
    assert pc["human"]["points"].shape[0] == B and pc["human"]["points"].shape[1] == pc["object"]["points"].shape[1]
    assert pc["object"]["centers"].shape == (B, 6) and pc["object"]["pca_axis"].shape == (B, 3, 3)
    pca_init = torch.linalg.qr(torch.randn(3, 3, generator=torch.Generator().manual_seed(1)))[0]
    torch.manual_seed(0)
    out = fit_recon_batch(fitter, gen, data, smpl_init, kpts, obj_points, pca_init=pca_init, silhouette=sil, mini_batch_size=2, max_iter=1, steps_per_iter=2)
    R = out["obj_R"]
    assert R.shape == (B, 3, 3) and float((R @ R.transpose(1, 2) - torch.eye(3, device=dev)).abs().max()) < 1e-5
    assert float(torch.linalg.det(R.double()).min()) > 0.999
    assert 2 <= len(out["hist_smpl"]) <= (3 + 1) * 2 and (15 + 30) * 2 < len(out["hist_obj"]) <= (10 + 15 + 1 + 30) * 2     # early stops allowed
    assert all(np.isfinite(out["hist_smpl"])) and all(np.isfinite(out["hist_obj"]))
    assert out["obj_t"].shape == (B, 3) and bool(torch.isfinite(out["smpl"].pose).all())
    # rotation handed over from HVOP-Net instead of the network's PCA axes
    out2 = fit_recon_batch(fitter, gen, data, smpl_init, kpts, obj_points, obj_rot_init=torch.eye(3)[None].repeat(B, 1, 1), silhouette=sil,
                           mini_batch_size=8, max_iter=1, steps_per_iter=1)
    assert out2["obj_R"].shape == (B, 3, 3)
    with pytest.raises(ValueError, match="pca_init"):
        fit_recon_batch(fitter, gen, data, smpl_init, kpts, obj_points, silhouette=sil, max_iter=1, steps_per_iter=1)
    k = scale_body_kpts(torch.tensor([[[1024.0, 768.0, 0.9]]], device=dev).repeat(1, 25, 1), torch.tensor([[1024.0, 768.0]], device=dev))
    assert torch.allclose(k[0, 0], torch.tensor([256.0, 256.0, 0.9], device=dev))
