"""Host-side logic that needs no GPU: mini-batch combination and key-point scaling of the fit_recon driver, the HVOP-Net clip plan / masks,
the 8-bit resize tables of the frame preparation."""
import numpy as np
import torch

from oracle import frameio_ref as FR
from oracle import infill_ref as IR


def test_combine_mini_batches_and_kpt_scaling():
    from vistracker_b200.recon_driver import combine_mini_batches, scale_body_kpts
    mk = lambda b, n: {t: {"points": torch.rand(b, n, 3), "parts": torch.randint(0, 14, (b, n)), "centers": torch.rand(b, 6),
                            "pca_axis": torch.rand(b, 3, 3), "visibility": torch.rand(b, 1)} for t in ("human", "object")}
    a, b = mk(2, 50), mk(1, 40)
    c = combine_mini_batches([a, b], 40)                                            # recon_fit_behave.py:152-183
    assert c["human"]["points"].shape == (3, 40, 3) and c["object"]["parts"].shape == (3, 40) and c["object"]["centers"].shape == (3, 6)
    assert torch.equal(c["human"]["points"][:2], a["human"]["points"][:, :40]) and torch.equal(c["object"]["pca_axis"][2:], b["object"]["pca_axis"])
    k = torch.tensor([[[1024.0, 768.0, 0.5], [424.0, 168.0, 1.0]]])
    out = scale_body_kpts(k, torch.tensor([[1024.0, 768.0]]))                       # recon_fit_base.py:397-409 with unit scales
    assert torch.allclose(out[0], torch.tensor([[256.0, 256.0, 0.5], [0.0, 0.0, 1.0]]))
    out2 = scale_body_kpts(k, torch.tensor([[1024.0, 768.0]]), resize_scale=torch.tensor([2.0]), crop_scale=torch.tensor([2.0]))
    assert torch.allclose(out2[0, 0, :2], torch.tensor([(2048.0 - 1024.0 + 1200.0) * 512 / 2400, (1536.0 - 768.0 + 1200.0) * 512 / 2400]))


def test_infill_clip_plan_and_masks_follow_the_reference_loop():
    from vistracker_b200.infill import CondMotionInfillAutoreg
    drv = CondMotionInfillAutoreg.__new__(CondMotionInfillAutoreg)
    drv.clip_len, drv.window, drv.init_thres = 180, 30, 0.5
    for L in (100, 180, 200, 400, 1500):
        plan = drv.clip_plan(L)
        ref = [(0, min(180, L), 0)] + [(i, min(180, L - i), min(30, L - i)) for i in range(0, L - 180 + 1 + 30, 30)]   # test_infill_autoreg.py:93,116-117
        assert plan == ref, L
        assert all(s + T <= L and T > 0 for s, T, _ in plan)
    occ = np.linspace(0, 1, 400).astype(np.float32)
    rows = drv._masks(occ, 0.3, drv.clip_plan(400))
    assert rows.shape == (len(drv.clip_plan(400)), 180)
    assert np.array_equal(rows[0].astype(bool), occ[:180] < 0.5)                     # first clip: init_thres, no context frames
    m2 = occ[30:210] < 0.3; m2[:30] = False
    assert np.array_equal(rows[2].astype(bool), m2)
    assert rows[-1, 160:].sum() == 0                                                 # the short last clip is padded with "visible"


def test_resize_tables_match_the_oracle_tables():
    from vistracker_b200.frameio import resize_table
    for d, s in ((512, 1200), (64, 150), (100, 100), (33, 200), (5, 3)):
        t = resize_table(d, s)
        i, a0, a1 = FR._coeffs(d, s)
        assert t.dtype == np.int32 and np.array_equal(t[0], i) and np.array_equal(t[1], a0) and np.array_equal(t[2], a1)
        assert int(t[0].max()) <= s - 1 and bool(((t[1] + t[2]) >= 2047).all()) and bool(((t[1] + t[2]) <= 2049).all())


def test_position_embedding_host_table_equals_the_oracle():
    from vistracker_b200.infill import position_embedding
    for L, D in ((180, 128), (180, 32), (160, 160), (47, 33), (1, 8)):
        # same operation order; ATen's vectorised / scalar-tail sin and cos may differ by one ulp depending on how the rows are split over threads
        assert float((position_embedding(L, D) - IR.position_embedding(L, D)).abs().max()) <= 2.4e-7


def test_triplane_view_transforms_match_reference_static_method():
    import os
    from vistracker_b200.render import TriplaneNrRenderer
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "render_views.npz"))
    pts = torch.from_numpy(g["points"])
    for view in ("right", "back", "top"):
        assert np.array_equal(TriplaneNrRenderer.transform_view(pts, view).numpy(), g[view]), view
    assert np.array_equal(TriplaneNrRenderer.transform_view(pts, "top", 5.0).numpy(), g["top_z5"])
    batched = TriplaneNrRenderer.transform_view(pts[None].repeat(2, 1, 1), "right")
    assert np.array_equal(batched[1].numpy(), g["right"])


def test_roi_setup_arithmetic_matches_reference_functions():
    """make_bbox_square / to_original_bbox / compute_K_roi / cvt_masks against the reference's (roi_small.npz); the mask crop (detectron2,
    restated on torchvision.ops.roi_align) is checked for its geometry."""
    import os
    from vistracker_b200.render import SilLossROI, cvt_masks, make_bbox_square, mask_bboxes_xyxy, roi_setup, to_original_bbox
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "roi_small.npz"))
    sq = make_bbox_square(g["boxes_xywh"], 0.3)
    assert np.allclose(sq, g["squares"], rtol=0, atol=1e-12)
    for i in range(6):
        o = to_original_bbox(sq[i], 1200 / 512, g["centers"][i], 1200)
        assert np.allclose(o, g["orig"][i], rtol=0, atol=1e-9)
        assert np.abs(SilLossROI.compute_K_roi(o).numpy() - g["Ks"][i]).max() < 1e-6
    K = SilLossROI.compute_K_roi(g["orig"][0], image_width=1920, fx=918.457763671875, fy=918.4373779296875, cx=956.9661865234375, cy=555.944580078125)
    assert np.abs(K.numpy() - g["K_icap"]).max() < 1e-6
    keep = torch.stack([cvt_masks(torch.from_numpy(p), torch.from_numpy(o)) for p, o in zip(g["ps"], g["ob"])])
    assert np.array_equal(keep.numpy(), g["keep"])
    # geometry of the crop: a 40 x 20 object box centred at (100, 60) in a 128 x 192 network input, person to its left overlapping it
    om = torch.zeros(2, 128, 192); om[:, 50:70, 80:120] = 1.0
    pm = torch.zeros(2, 128, 192); pm[:, 40:90, 60:95] = 1.0
    assert mask_bboxes_xyxy(om).tolist() == [[80, 50, 120, 70]] * 2
    keep, ref, K = roi_setup(pm, om, torch.tensor([[1024.0, 768.0], [900.0, 700.0]]), rend_size=64, net_input_size=128, crop_size=1200)
    assert keep.shape == (2, 64, 64) and ref.shape == (2, 64, 64) and K.shape == (2, 3, 3)
    # the square is 52 px wide (40 * 1.3) around (100, 60): the object spans 40/52 of the width and 20/52 of the height of the crop
    cols, rows = ref[0].any(0).nonzero().flatten(), ref[0].any(1).nonzero().flatten()
    assert abs((int(cols[-1]) - int(cols[0]) + 1) - 64 * 40 / 52) <= 1.5 and abs((int(rows[-1]) - int(rows[0]) + 1) - 64 * 20 / 52) <= 1.5
    assert abs((int(cols[0]) + int(cols[-1])) / 2 - 31.5) <= 1.0 and abs((int(rows[0]) + int(rows[-1])) / 2 - 31.5) <= 1.0
    assert bool((keep[0][ref[0] > 0] == 1).all()) and float(keep[0].min()) == 0.0          # person-only pixels are ignored, object pixels kept
    side = 52 * 1200 / 128
    assert abs(float(K[0, 0, 0]) - 979.7844 / side) < 1e-6 and abs(float(K[1, 0, 2]) - (1018.952 - (900 - 600 + (100 - 26) * 1200 / 128)) / side) < 1e-5


def test_driver_arithmetic_matches_reference_methods():
    """scale_body_kpts and combine_mini_batches against the reference's own methods (tests/golden/driver_small.npz)."""
    import os
    from vistracker_b200.recon_driver import combine_mini_batches, scale_body_kpts
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "driver_small.npz"))
    t = lambda k: torch.from_numpy(g[k])
    got = scale_body_kpts(t("kpts"), t("crop_center"), resize_scale=t("resize_scale"), crop_scale=t("crop_scale"))
    assert np.allclose(got.numpy(), g["scaled"], rtol=1e-6, atol=1e-4)
    assert np.allclose(scale_body_kpts(t("kpts"), t("crop_center")).numpy(), g["scaled_unit"], rtol=1e-6, atol=1e-4)
    pcs = [{tt: {k: t(f"pc{i}.{tt}.{k}") for k in ("points", "parts", "centers", "pca_axis", "visibility")} for tt in ("human", "object")} for i in range(3)]
    comb = combine_mini_batches(pcs, 40)
    for tt in ("human", "object"):
        for k in ("points", "parts", "centers", "pca_axis", "visibility"):
            assert tuple(comb[tt][k].shape) == g[f"comb.{tt}.{k}"].shape and np.array_equal(comb[tt][k].numpy(), g[f"comb.{tt}.{k}"]), (tt, k)


def test_compute_pca_matches_sklearn_conventions():
    """geom.compute_pca: sign='v' reproduces PCAUtil.compute_pca as the reference computes it with the installed scikit-learn (golden),
    sign='u' the pre-1.5 svd_flip rule; both span the same axes."""
    import os
    from sklearn.utils.extmath import svd_flip
    from vistracker_b200.geom import compute_pca
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "infill_small.npz"))
    X = g["pca_template"]
    v = compute_pca(X, sign="v")
    assert np.abs(v - g["pca_components_installed_sklearn"]).max() < 1e-10
    U, S, Vt = np.linalg.svd(X - X.mean(0), full_matrices=False)
    _, Vu = svd_flip(U, Vt, u_based_decision=True)
    u = compute_pca(torch.from_numpy(X), sign="u")
    assert np.abs(u - Vu).max() < 1e-10
    assert np.abs(np.abs(u @ v.T) - np.eye(3)).max() < 1e-9 and np.abs(u @ u.T - np.eye(3)).max() < 1e-10
    assert compute_pca(X).tolist() == u.tolist()                                   # default: the convention the checkpoints were trained with


def test_contact_pairs_equal_the_reference_double_loop():
    """ReconFitterTriVisFull.contact_pairs (vectorised sorts / counts) against the literal frame x part loop of compute_contact_loss
    (recon/recon_fit_trivis_full.py:405-449) written out in Python: same clouds, same order, incl. a frame without human contacts, a frame
    without object contacts and parts present on one side only."""
    from types import SimpleNamespace
    from vistracker_b200.recon_fit import ReconFitterTriVisFull
    rng = np.random.default_rng(5)
    B, Nh, No = 5, 300, 90
    labels = torch.from_numpy(rng.integers(0, 14, Nh))
    df_hum_o = torch.from_numpy(rng.uniform(0, 0.3, (B, Nh)).astype(np.float32))
    df_obj_h = torch.from_numpy(rng.uniform(0, 0.3, (B, No)).astype(np.float32))
    part_o = torch.from_numpy(rng.standard_normal((B, 14, No)).astype(np.float32))
    part_o[:, 5] = -10.0                                 # part 5 never predicted on the object
    df_hum_o[1] = 1.0                                    # no human contact in frame 1
    df_obj_h[3] = 1.0                                    # no object contact in frame 3
    shim = SimpleNamespace(part_labels=labels)
    h_idx, o_idx, h_off, o_off = ReconFitterTriVisFull.contact_pairs(shim, df_hum_o, df_obj_h, part_o)
    # literal restatement of the reference loop, as lists of flat indices
    hs, os_ = [], []
    po = part_o.argmax(1)
    for b in range(B):
        mh, mo = df_hum_o[b] < 0.08, df_obj_h[b] < 0.08
        if int(mh.sum()) + int(mo.sum()) == 0 or int(mo.sum()) == 0 or int(mh.sum()) == 0:
            continue
        hv, ov = torch.where(mh)[0], torch.where(mo)[0]
        lh, lo = labels[hv], po[b][ov]
        for i in range(14):
            if i not in lh or i not in lo:
                continue
            hs.append(hv[lh == i] + b * Nh); os_.append(ov[lo == i] + b * No)
    assert len(hs) == h_off.numel() - 1 == o_off.numel() - 1 and len(hs) > 10
    for n, (a, c) in enumerate(zip(hs, os_)):
        assert torch.equal(h_idx[h_off[n]:h_off[n + 1]], a) and torch.equal(o_idx[o_off[n]:o_off[n + 1]], c)
    assert h_off.dtype == torch.int32 and int(h_off[-1]) == h_idx.numel() and int(o_off[-1]) == o_idx.numel()
    # nothing in contact anywhere
    assert ReconFitterTriVisFull.contact_pairs(shim, df_hum_o + 1, df_obj_h, part_o) is None


def test_phase_schedules_follow_the_reference_chains():
    from vistracker_b200.recon_fit import ReconFitterTriVisFull as F
    s = F.smpl_phase_schedule(1, 1, 1, 100)              # the tri-vis driver's call (recon_fit_triplane.py:66)
    assert len(s) == 103 and s[0] == ("global", False) and s[1] == ("smpl all pose", True) and s[2] == ("kpts", False) and s[-1] == ("kpts", False)
    s = F.smpl_phase_schedule(10, 10, 5, 100)            # the defaults of recon_fit_behave.py:393-395
    assert [p for p, _ in s[:10]] == ["global"] * 10 and s[10] == ("smpl all pose", True) and s[19][0] == "smpl all pose" and s[20][0] == "kpts"
    s = F.smpl_phase_schedule(1, 0, 1, 3)                # iter_for_pose = 0: the `elif` chain never reaches 'kpts'
    assert [p for p, _ in s] == ["global"] + ["smpl all pose"] * 4
    o = F.object_phase_schedule(15, 30, 10, 100)
    assert len(o) == 155 and o[14] == ("object only", False, 1) and o[15] == ("sil", True, 1) and o[44] == ("sil", False, 30)
    assert o[45][:2] == ("joint", True) and abs(o[45][2] - 31 / 3) < 1e-12 and o[46][:2] == ("joint", False)
    o = F.object_phase_schedule(2, 0, 1, 3)              # no silhouette phase: `it == it_obj and it != it_obj + it_sil` is False, joint starts at it_obj
    assert [p for p, _, _ in o] == ["object only"] * 2 + ["joint"] * 4 and o[2][1]


def test_generator_random_stream_trick_reproduces_the_reference_draw_order():
    """gen_pc_batch draws the resampling noise BEFORE the survivor counts are known (generator.py of this package): a dummy randint advances
    torch's CPU generator exactly like the reference's ``randint(high, (1, n))`` whatever ``high`` is (one 32-bit draw per element below 2^32),
    the real index draw is made later from the generator state saved in front of it, and the state after the round is restored.  The tensors
    and the generator state afterwards must equal the reference's sequential order (recon/gen/generator.py:185-205: per frame randint, randn)."""
    import torch
    n, highs = 4000, [17, 19999, 30000, 2, 123456, 1 << 20]

    def reference_order():
        torch.manual_seed(11)
        out = [(torch.randint(h, (1, n)), torch.randn(1, n, 3)) for h in highs]
        return out, torch.rand(5)

    def early_noise_order():
        torch.manual_seed(11)
        states, noise = [], []
        for _ in highs:                                   # before the counts are known
            states.append(torch.get_rng_state())
            torch.randint(2, (1, n))
            noise.append(torch.randn(1, n, 3))
        final = torch.get_rng_state()
        idx = []
        for s, h in zip(states, highs):                   # after: the real ranges
            torch.set_rng_state(s)
            idx.append(torch.randint(h, (1, n)))
        torch.set_rng_state(final)
        return list(zip(idx, noise)), torch.rand(5)

    a, ta = reference_order()
    b, tb = early_noise_order()
    assert all(torch.equal(x[0], y[0]) and torch.equal(x[1], y[1]) for x, y in zip(a, b))
    assert torch.equal(ta, tb)                            # the generator continues as in the reference


def test_human_shaped_synthetic_body_model_has_the_structure_of_smplh():
    """synth_smpl.synthetic_smplh_surface (the body of the bench's C4 batch): SMPL-H buffer shapes, a regressor and skinning weights with rows
    summing to one and 4 joints per vertex, vertices that are neighbours in index order, and -- posed by the bench's own motion through the
    CPU restatement of SMPL_Layer.forward -- a body that stays inside the triplane around the batch's body centre."""
    from oracle.smpl_ref import smpl_forward
    from vistracker_b200.synth import synthetic_recon_batch
    from vistracker_b200.synth_smpl import NUM_JOINTS, NUM_VERTS, synthetic_smplh, synthetic_smplh_surface
    m, ref = synthetic_smplh_surface(seed=3), synthetic_smplh(seed=3)
    assert set(m) == set(ref)
    for k in ref:
        if torch.is_tensor(ref[k]):
            assert m[k].shape == ref[k].shape and m[k].dtype == ref[k].dtype, k
    assert m["parents"] == ref["parents"]
    assert torch.allclose(m["th_J_regressor"].sum(1), torch.ones(NUM_JOINTS), atol=1e-5)
    assert torch.allclose(m["th_weights"].sum(1), torch.ones(NUM_VERTS), atol=1e-5) and int((m["th_weights"] > 0).sum(1).max()) == 4
    assert int(m["th_faces"].max()) == NUM_VERTS - 1 and int(m["th_faces"].min()) == 0
    h = synthetic_recon_batch(8, size=16, seed=4)
    verts, jtr = smpl_forward(m, h["pose"], h["betas"], h["trans"])[:2]
    assert torch.isfinite(verts).all()
    c = verts - h["body_center"][:, None]
    assert float(c.abs().max()) < 1.0                                             # every tap of every vertex inside the three triplane views
    assert float((jtr[:, 0] - h["body_center"]).abs().max()) < 0.1               # the root joint rides on the body centre
    step = (verts[:, 1:] - verts[:, :-1]).norm(dim=-1).median()
    assert float(step) < 0.05                                                    # consecutive vertices are surface neighbours (0.5 m for the cloud)
