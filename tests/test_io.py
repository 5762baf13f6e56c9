"""The on-disk formats of vistracker_b200/io.py: keys, shapes and dtypes as the reference's writers produce them
(preprocess/fit_SMPLH_kpts.py:250-261, recon/opt_utils.py:134-141, recon/recon_fit_base.py:278-313,830-844)."""
import os
import pickle as pkl

import numpy as np
import pytest
import torch

from vistracker_b200 import io as vio


def test_output_folders_and_param_files(tmp_path):
    paths = [f"/data/Date03_Sub03_chairwood_hand/t{i:04d}.000/k1.color.jpg" for i in range(3)]
    folders = vio.output_folders(str(tmp_path), paths, "test-release")
    assert folders[1] == os.path.join(str(tmp_path), "Date03_Sub03_chairwood_hand", "t0001.000", "test-release") and os.path.isdir(folders[2])
    pose, betas, trans = torch.randn(3, 156), torch.randn(3, 10), torch.randn(3, 3)
    files = vio.save_smpl_params(folders, 1, pose, betas, trans)
    d = pkl.load(open(files[2], "rb"))
    assert sorted(d) == ["betas", "pose", "score", "trans"] and d["pose"].shape == (156,) and d["pose"].dtype == np.float32
    assert np.array_equal(d["trans"], trans[2].numpy()) and float(d["score"]) == 0.0
    R = torch.linalg.qr(torch.randn(3, 3, 3))[0]
    R = R * torch.sign(torch.linalg.det(R))[:, None, None]
    ofiles = vio.save_object_params(folders, 1, R, torch.randn(3, 3), torch.ones(3))
    o = pkl.load(open(ofiles[0], "rb"))
    assert sorted(o) == ["rot", "scale", "trans"] and o["rot"].shape == (3, 3) and o["scale"].shape == ()


def test_smplt_fit_files_round_trip_and_skip(tmp_path):
    poses, betas, trans = np.random.rand(4, 156).astype(np.float32), np.random.rand(4, 10).astype(np.float32), np.random.rand(4, 3).astype(np.float32)
    files = [str(tmp_path / f"f{i}.k1.smplfit_temporal.pkl") for i in range(4)]
    n = vio.save_smplt_fits(files, torch.from_numpy(poses), betas, trans, skip=[False, True, False, False])
    assert n == 3 and not os.path.exists(files[1])
    d = pkl.load(open(files[0], "rb"))
    assert sorted(d) == ["betas", "pose", "trans"]
    P, B, T = vio.load_smplt_fits([files[0], files[2], files[3]])
    assert np.array_equal(P, poses[[0, 2, 3]]) and np.array_equal(T, trans[[0, 2, 3]]) and B.shape == (3, 10)


def test_neural_recon_npz(tmp_path):
    folders = [str(tmp_path / f"fr{i}") for i in range(2)]
    for f in folders:
        os.makedirs(f)
    batch = {t: {"points": torch.randn(2, 50, 3), "pca_axis": torch.randn(2, 3, 3), "parts": torch.randint(0, 14, (2, 50)),
                 "centers": torch.randn(2, 6), "visibility": torch.rand(2, 1)} for t in ("human", "object")}
    files = vio.save_neural_recon(folders, 1, batch)
    assert os.path.basename(files[0]) == "k1_densepc.npz"
    z = np.load(files[1], allow_pickle=True)
    assert sorted(z.files) == ["human", "object"]
    h = z["human"].item()
    assert sorted(h) == ["centers", "parts", "pca_axis", "points", "visibility"] and h["points"].shape == (50, 3)
    assert np.array_equal(h["parts"], batch["human"]["parts"][1].numpy())


def test_smplh_pose_padding_and_param_copy():
    """a14: SMPLHGenerator.get_smplh's 72 -> 156 padding and ReconFitterBase.copy_smpl_params on host containers."""
    import types
    from vistracker_b200.recon_fit import SMPLParams, copy_smpl_params, smplh_pose
    hm = np.arange(90, dtype=np.float32) * 0.01
    p72 = np.random.default_rng(0).standard_normal((4, 72)).astype(np.float32)
    out = smplh_pose(p72, hm)
    assert out.shape == (4, 156) and np.array_equal(out[:, :66].numpy(), p72[:, :66]) and np.array_equal(out[:, 66:].numpy(), np.tile(hm, (4, 1)))
    p156 = np.random.default_rng(1).standard_normal((4, 156)).astype(np.float32)
    assert np.array_equal(smplh_pose(p156).numpy(), p156)
    with pytest.raises(ValueError, match="mean hand pose"):
        smplh_pose(p72)
    layer = types.SimpleNamespace(device=torch.device("cpu"), th_faces=None)
    a = SMPLParams(layer, None, torch.from_numpy(p156), torch.zeros(4, 10), torch.zeros(4, 3))
    b = SMPLParams(layer, None, torch.from_numpy(p156) + 1, torch.ones(4, 10), torch.ones(4, 3))
    copy_smpl_params(b, a)
    assert torch.equal(a.pose, b.pose) and torch.equal(a.trans, b.trans) and torch.equal(a.betas[:, :2], b.betas[:, :2])
    assert float(a.betas[:, 2:].abs().max()) == 0.0                               # the other betas are not copied
    c = SMPLParams.from_smpl(b)
    assert torch.equal(c.pose.detach(), b.pose.detach()) and c.global_pose.requires_grad and c.global_pose is not b.global_pose


def test_sequence_packs_roundtrip(tmp_path):
    T = 7
    rng = np.random.default_rng(2)
    frames = [f"t{i:04d}.000" for i in range(T)]
    poses, betas, trans = rng.standard_normal((T, 156)).astype(np.float32), rng.standard_normal((T, 10)).astype(np.float32), rng.standard_normal((T, 3)).astype(np.float32)
    f = vio.pack_smplt(str(tmp_path / "smplt" / "seq_k1.pkl"), frames, "male", torch.from_numpy(poses), betas, trans)
    d = vio.load_packed(f)
    assert list(d) == ["poses", "betas", "trans", "obj_angles", "obj_trans", "obj_scales", "gender", "frames"]          # pack_smplt.py:45-63
    assert np.array_equal(d["poses"], poses) and d["obj_angles"].shape == (T, 3, 3) and d["gender"] == "male" and d["frames"] == frames
    pca, nt, vis = rng.standard_normal((T, 3, 3)).astype(np.float32), rng.standard_normal((T, 3)).astype(np.float32), rng.random((T, 1)).astype(np.float32)
    f = vio.pack_recon(str(tmp_path / "recon_x" / "seq_k1.pkl"), frames, "female", "x", pca, nt, vis)
    import joblib
    raw = joblib.load(f)
    assert list(raw) == ["neural_pca", "neural_trans", "recon_exist", "neural_visibility", "recon_name", "frames", "gender"]
    assert isinstance(raw["neural_pca"], list) and raw["neural_pca"][0].shape == (3, 3) and raw["neural_visibility"][0].shape == (1,)
    assert np.array(raw["neural_visibility"])[:, 0].shape == (T,)                  # how test_infill_autoreg.py:80 reads it
    d = vio.load_packed(f)
    assert np.array_equal(d["neural_pca"], pca) and bool(d["recon_exist"].all())
    ang = rng.standard_normal((T, 3, 3)).astype(np.float32)
    f = vio.pack_recon(str(tmp_path / "recon_y" / "seq_k1.pkl"), frames, "male", "y", pca, nt, vis, poses, betas, trans, trans * 2, ang, trans + 1, np.ones(T))
    d = vio.load_packed(f)
    assert d["obj_angles"].shape == (T, 3, 3) and np.array_equal(d["root_joints"], trans * 2) and d["obj_scales"].shape == (T,) and "poses" in d


def test_triplane_png_roundtrip(tmp_path):
    rng = np.random.default_rng(3)
    masks = torch.from_numpy((rng.random((2, 3, 32, 32)) > 0.6).astype(np.uint8))
    files = [str(tmp_path / f"t{i:04d}.000" / "k1.smooth_triplane.png") for i in range(2)]
    assert vio.save_triplane_png(files, masks) == files
    back = vio.load_triplane_png(files[1])
    assert back.shape == (32, 32, 3) and back.dtype == np.uint8 and set(np.unique(back)) <= {0, 255}
    assert np.array_equal(back.transpose(2, 0, 1) // 255, masks[1].numpy())           # channel 0 = right, 1 = back, 2 = top
    with pytest.raises(ValueError, match="expected masks"):
        vio.save_triplane_png(files, masks[:, :2])


def test_packed_batch_slices_by_frame_name(tmp_path):
    T = 6
    rng = np.random.default_rng(5)
    frames = [f"t{i:04d}.{(i * 33) % 1000:03d}" for i in range(T)]
    pca, nt, vis = rng.standard_normal((T, 3, 3)), rng.standard_normal((T, 3)), rng.random((T, 1))
    poses, betas, trans, ang = rng.standard_normal((T, 156)), rng.standard_normal((T, 10)), rng.standard_normal((T, 3)), rng.standard_normal((T, 3, 3))
    f = vio.pack_recon(str(tmp_path / "recon_z" / "seq_k1.pkl"), frames, "male", "z", pca, nt, vis, poses, betas, trans, trans, ang, trans, np.ones(T))
    import joblib
    raw = joblib.load(f)                                                          # lists, as the reference stores the neural entries
    paths = [f"/data/Date03_Sub03_chairwood_hand/{frames[i]}/k1.color.jpg" for i in (4, 1, 2)]
    b = vio.packed_batch(raw, paths)
    assert b["frame_inds"].tolist() == [4, 1, 2] and np.array_equal(b["poses"], poses[[4, 1, 2]]) and np.array_equal(b["obj_angles"], ang[[4, 1, 2]])
    assert np.array_equal(b["neural_pca"], pca[[4, 1, 2]]) and np.allclose(b["occ_ratios"], vis[[4, 1, 2], 0])
    with pytest.raises(AssertionError, match="kinect id"):
        vio.packed_batch(raw, [paths[0].replace("k1.", "k2.")])
    with pytest.raises(ValueError):
        vio.packed_batch(raw, ["/data/seq/t9999.000/k1.color.jpg"])


def test_ply_roundtrip_binary_and_ascii(tmp_path):
    rng = np.random.default_rng(8)
    v = rng.standard_normal((50, 3)).astype(np.float32)
    f = rng.integers(0, 50, (80, 3))
    p = vio.save_ply(str(tmp_path / "seq" / "t0001.000" / "k1.smplfit_smoothed.ply"), torch.from_numpy(v), f)
    v2, f2 = vio.load_ply(p)
    assert np.array_equal(v2.astype(np.float32), v) and np.array_equal(f2, f) and f2.dtype == np.int64
    head = open(p, "rb").read(200).decode("ascii", "ignore")
    assert head.startswith("ply\nformat binary_little_endian 1.0\nelement vertex 50\n") and "property list uchar int vertex_indices" in head
    asc = tmp_path / "a.ply"
    asc.write_text("ply\nformat ascii 1.0\ncomment made by hand\nelement vertex 3\nproperty float x\nproperty float y\nproperty float z\n"
                   "property uchar red\nelement face 1\nproperty list uchar int vertex_indices\nend_header\n0 0 0 255\n1 0 0 255\n0 1.5 0 255\n3 0 1 2\n")
    v3, f3 = vio.load_ply(str(asc))
    assert v3.tolist() == [[0, 0, 0], [1, 0, 0], [0, 1.5, 0]] and f3.tolist() == [[0, 1, 2]]


def test_writers_match_what_the_reference_writes(tmp_path):
    """tests/golden/io_formats.npz holds what the reference's own writers put on disk for these inputs (save_neural_recon, save_outputs with
    save_smplfits, BaseFitter.save_results): same file names, key order, dtypes, shapes, values."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "io_formats.npz"), allow_pickle=False)
    paths = [str(p) for p in g["paths"]]
    B = len(paths)
    # ---- k{tid}_densepc.npz
    folders = vio.output_folders(str(tmp_path / "recon"), paths, "test-release")
    recon_batch = {t: {k: torch.from_numpy(g[f"in.{t}.{k}"]) for k in ("points", "pca_axis", "parts", "centers", "visibility")} for t in ("human", "object")}
    files = vio.save_neural_recon(folders, 1, recon_batch)
    for i, f in enumerate(files):
        assert f.endswith(os.path.join("Date03_Sub03_chairwood_hand", os.path.basename(os.path.dirname(paths[i])), "test-release", "k1_densepc.npz"))
        d = np.load(f, allow_pickle=True)
        order = [f"{t}.{k}" for t in d.files for k in d[t].item()]
        assert order == [str(x) for x in g[f"densepc{i}.order"]]
        for t in d.files:
            for k, v in d[t].item().items():
                ref = g[f"densepc{i}.{t}.{k}"]
                assert v.dtype == ref.dtype and v.shape == ref.shape and np.array_equal(v, ref, equal_nan=True), (t, k)
    # ---- k{tid}.smpl.pkl / k{tid}.object.pkl
    folders = vio.output_folders(str(tmp_path / "recon"), paths, "test-releasev2")
    vio.save_smpl_params(folders, 1, g["in.pose"], g["in.betas"], g["in.trans"])
    from oracle.geom_ref import project_so3
    R = project_so3(torch.from_numpy(g["in.obj_R"]))                                # a host array is taken as projected already (see io.py)
    vio.save_object_params(folders, 1, R, g["in.obj_t"], g["in.obj_s"])
    for i, folder in enumerate(folders):
        sm = pkl.load(open(os.path.join(folder, "k1.smpl.pkl"), "rb"))
        ob = pkl.load(open(os.path.join(folder, "k1.object.pkl"), "rb"))
        assert list(sm) == [str(x) for x in g[f"smpl{i}.order"]] and list(ob) == [str(x) for x in g[f"object{i}.order"]]
        for k, v in sm.items():
            ref = g[f"smpl{i}.{k}"]
            assert np.asarray(v).dtype == ref.dtype and np.array_equal(np.asarray(v), ref), k
        for k, v in ob.items():
            ref = g[f"object{i}.{k}"]
            assert np.asarray(v).shape == ref.shape and np.asarray(v).dtype == ref.dtype, k
            assert np.allclose(np.asarray(v), ref, atol=2e-6), k
    # ---- k{kid}.smplfit_temporal.pkl, frames without confident key points skipped
    outfiles = [str(tmp_path / "seq" / os.path.basename(os.path.dirname(p)) / "k1.smplfit_temporal.pkl") for p in paths]
    skip = torch.from_numpy(g["in.scores"]).sum(1) < 0.1                            # BaseFitter.skip_frame
    assert vio.save_smplt_fits(outfiles, g["in.pose"], g["in.betas"], g["in.trans"], skip=skip) == int(g["smplt.written"].sum())
    assert [os.path.isfile(f) for f in outfiles] == g["smplt.written"].tolist()
    d0 = pkl.load(open(outfiles[0], "rb"))
    assert list(d0) == [str(x) for x in g["smplt.order"]]
    for k, v in d0.items():
        assert v.dtype == g[f"smplt0.{k}"].dtype and np.array_equal(v, g[f"smplt0.{k}"])


def test_sequence_packs_match_the_reference_packers(tmp_path):
    """pack_formats.npz: what the reference's preprocess/pack_recon.py (neural-only and full) and pack_smplt.py wrote after READING the
    per-frame files this package's writers produced.  pack_recon / pack_smplt build the same packs straight from the arrays."""
    import joblib
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "pack_formats.npz"))
    frames = [str(f) for f in g["frames"]]
    T = len(frames)

    def check(tag, d):
        assert list(d) == [str(k) for k in g[f"{tag}.order"]], tag
        for k, v in d.items():
            ref = g[f"{tag}.{k}"]
            if f"{tag}.{k}.islist" in g.files:
                assert isinstance(v, list) and len(v) == T, (tag, k)
                v = np.stack([np.asarray(x) for x in v], 0)
            if ref.dtype.kind in "US":
                assert np.array_equal(np.asarray(v).astype(str), ref.astype(str)), (tag, k)
            else:
                assert np.asarray(v).shape == ref.shape, (tag, k, np.asarray(v).shape, ref.shape)
                assert np.allclose(np.asarray(v, dtype=np.float64), ref.astype(np.float64), atol=2e-6, equal_nan=True), (tag, k)

    pca, rel, vis = g["in.pca"], g["in.centers"][:, 3:], g["in.vis"]
    f = vio.pack_recon(str(tmp_path / "a.pkl"), frames, "male", "test-release", pca, rel, vis)
    check("neural", joblib.load(f))
    root = g["in.trans"] + 0.125                                                   # the stand-in of get_root_joint used when the golden was made
    f = vio.pack_recon(str(tmp_path / "b.pkl"), frames, "male", "test-releasev2", pca, rel, vis, g["in.pose"], g["in.betas"], g["in.trans"], root,
                       g["in.rot"], g["in.obj_t"], g["in.obj_s"])
    check("full", joblib.load(f))
    f = vio.pack_smplt(str(tmp_path / "c.pkl"), frames, "male", g["in.pose"], g["in.betas"], g["in.trans"])
    check("smplt", joblib.load(f))


def test_infill_output_pack_matches_reference_save_output(tmp_path):
    import joblib
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "infill_io.npz"))
    dat = {k[3:]: (g[k].tolist() if k == "in.frames" else (str(g[k]) if k == "in.gender" else g[k])) for k in g.files if k.startswith("in.")}
    dat = {k: dat[k] for k in ("poses", "betas", "trans", "obj_angles", "obj_trans", "obj_scales", "gender", "frames")}
    # the reference receives rot_pred (real rotations) and stores the transpose; the in-filler here already returns the stored layout
    f = vio.save_infill_output(dat, str(tmp_path / "a" / "seq_k1.pkl"), torch.from_numpy(g["rot_pred"]).transpose(1, 2), torch.from_numpy(g["trans_pred"]))
    for tag, path in (("filled", f), ("orig", vio.save_infill_output(dat, str(tmp_path / "b" / "seq_k1.pkl")))):
        d = joblib.load(path)
        assert list(d) == [str(k) for k in g[f"{tag}.order"]], tag
        for k, v in d.items():
            ref = g[f"{tag}.{k}"]
            if ref.dtype.kind in "US":
                assert np.array_equal(np.asarray(v).astype(str), ref.astype(str)), (tag, k)
            else:
                assert np.asarray(v).shape == ref.shape and np.allclose(np.asarray(v, dtype=np.float64), ref.astype(np.float64)), (tag, k)
