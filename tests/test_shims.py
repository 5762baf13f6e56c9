"""Entry-point shims (vistracker_b200/shims): the reference's module paths resolve, the argparse front ends accept the exact command lines of
scripts/demo.sh, ``merge_configs`` fills the fields ``fit_recon`` reads, and the sequence reader follows the BEHAVE file conventions -- CPU only."""
import json
import os
import shlex

import numpy as np
import pytest

from vistracker_b200 import shims
from vistracker_b200.shims import fit_SMPLH_30fps, fit_SMPLH_smoothed, recon_fit_trivis_full, render_triplane_nr, seqio

SEQ = "/data/behave/Date03_Sub03_chairwood_hand"
# scripts/demo.sh lines 14, 19, 23, 26, 36 with ${seq} substituted (the other lines run SmoothNet / HVOP-Net / packing / visualisation)
DEMO = {
    "preprocess/fit_SMPLH_30fps.py": f"-s {SEQ} -bs 512",
    "preprocess/fit_SMPLH_smoothed.py": f"-sn smplt-smoothed -s {SEQ}",
    "render/render_triplane_nr.py": f"-s {SEQ}",
    "recon/recon_fit_trivis_full.py#4": f"tri-vis-l2 -sn test-release -or neural -sr smplt-smoothed-fit -t 1 -bs 64 -tt smooth -neural_only -s {SEQ}",
    "recon/recon_fit_trivis_full.py#6": f"tri-vis-l2 -sr smplt-smoothed-fit -or smooth-hvopnet -sn test-releasev2 -s {SEQ}",
}


def test_reference_module_paths_resolve_to_the_shims():
    shims.install(force=True)
    import preprocess.fit_SMPLH_30fps as a
    import preprocess.fit_SMPLH_smoothed as b
    import recon.recon_fit_trivis_full as c
    import render.render_triplane_nr as d
    assert callable(a.main) and callable(a.SMPLHFitter30fps.fit_seq) and callable(b.main) and callable(d.main)
    assert callable(c.recon_fit) and callable(c.ReconFitterTriVisFull.get_parser) and callable(c.ReconFitterTriVisFull.merge_configs)
    assert callable(c.ReconFitterTriVisFull.fit_recon)


def test_parsers_accept_the_demo_sh_command_lines():
    a = fit_SMPLH_30fps.get_parser().parse_args(shlex.split(DEMO["preprocess/fit_SMPLH_30fps.py"]))
    assert (a.seq_folder, a.batch_size, a.kid, a.start, a.end, a.redo, a.init_type) == (SEQ, 512, 1, 0, None, False, "mocap")
    a = fit_SMPLH_smoothed.get_parser().parse_args(shlex.split(DEMO["preprocess/fit_SMPLH_smoothed.py"]))
    assert (a.seq_folder, a.smoothed_name, a.batch_size, a.kid) == (SEQ, "smplt-smoothed", 512, 1)
    a = render_triplane_nr.get_parser().parse_args(shlex.split(DEMO["render/render_triplane_nr.py"]))
    assert (a.seq_folder, a.kids, a.mesh_type, a.redo) == (SEQ, [1], "smooth", False)
    P = recon_fit_trivis_full.ReconFitterTriVisFull.get_parser()
    a = P.parse_args(shlex.split(DEMO["recon/recon_fit_trivis_full.py#4"]))
    assert (a.exp_name, a.save_name, a.obj_recon_name, a.smpl_recon_name, a.tid, a.batch_size, a.triplane_type, a.neural_only) == \
        ("tri-vis-l2", "test-release", "neural", "smplt-smoothed-fit", 1, 64, "smooth", True)
    a = P.parse_args(shlex.split(DEMO["recon/recon_fit_trivis_full.py#6"]))
    assert (a.save_name, a.obj_recon_name, a.batch_size, a.neural_only, a.filter_val, a.sparse_thres, a.redo) == \
        ("test-releasev2", "smooth-hvopnet", 96, False, 0.004, 0.03, False)
    with pytest.raises(SystemExit):                      # -sn / -sr / -or are required, as in the reference
        P.parse_args(["tri-vis-l2", "-s", SEQ])


def test_merge_configs_fills_what_fit_recon_reads():
    P = recon_fit_trivis_full.ReconFitterTriVisFull
    a = P.get_parser().parse_args(shlex.split(DEMO["recon/recon_fit_trivis_full.py#6"]) + ["-fs", "30", "-fe", "90", "-redo"])
    c = P.merge_configs(a, recon_fit_trivis_full.load_configs(a.exp_name))
    assert (c.batch_size, c.test_kid, c.save_name, c.seq_folder, c.start, c.end, c.redo, c.neural_only, c.pred_occ) == (96, 1, "test-releasev2", SEQ, 30, 90, True, False, True)
    assert (c.smpl_recon_name, c.obj_recon_name, c.triplane_type, c.filter_val) == ("smplt-smoothed-fit", "smooth-hvopnet", "smooth", 0.004)
    assert c.z_feat == "smpl-triplane" and c.exp_name == "tri-vis-l2"          # the network options survive the merge


def test_sequence_reader_follows_the_behave_layout(tmp_path):
    from PIL import Image
    seq = tmp_path / "Date03_Sub03_chairwood_synth"
    for t in ("t0002.000", "t0001.000", "notaframe"):
        (seq / t).mkdir(parents=True)
    (seq / "info.json").write_text(json.dumps({"cat": "chairwood", "gender": "female", "config": None, "intrinsic": None, "empty": None, "beta": None,
                                               "kinects": [0, 1, 2, 3]}))
    f = seq / "t0001.000"
    m = np.zeros((40, 60), np.uint8); m[10:20, 5:30] = 255
    Image.fromarray(m).save(f / "k1.person_mask.jpg", quality=95)          # only the jpg exists: the png name is tried first
    Image.fromarray(m).save(f / "k1.obj_mask.png")                          # falls back from obj_rend_mask.* to obj_mask.*
    (f / "k1.mocap.json").write_text(json.dumps({"pose": list(range(72)), "betas": [0.5] * 10}))
    (f / "k1.color.json").write_text(json.dumps({"body_joints": [100.0, 200.0, 0.9, 5.0, 6.0, 0.05] + [0.0] * 69}))
    r = seqio.FrameDataReader(str(seq))
    assert r.frames == ["t0001.000", "t0002.000"] and r.seq_name == seq.name and len(r) == 2
    assert r.cvt_end(None) == 2 and r.cvt_end(1) == 1 and r.cvt_end(9) == 2
    assert r.seq_info.get_gender() == "female" and r.seq_info.get_obj_name() == "chairwood"
    assert r.get_mask_file(0, 1, "person").endswith("k1.person_mask.jpg") and r.get_mask_file(0, 1, "obj").endswith("k1.obj_mask.png")
    assert r.get_mask(0, 1, "person").sum() > 200 and r.get_mask(1, 1, "person") is None
    p, b = r.get_mocap_params(0, 1)
    assert p.shape == (72,) and b.shape == (10,) and r.get_mocap_params(1, 1) == (None, None)
    k = r.get_body_kpts(0, 1, tol=0.1)
    assert k.shape == (25, 3) and k[0].tolist() == [100.0, 200.0, 0.9] and k[1, 2] == 0.0        # low-confidence joints are zeroed
    person, obj = seqio.load_masks(str(f / "k1.color.jpg"))
    assert person.shape == (40, 60) and obj.dtype == np.uint8


def test_skip_if_done_checks_the_files_the_reference_checks(tmp_path):
    """is_done (recon_fit_base.py:260-276): k<tid>_densepc.npz for -neural_only, the two parameter pickles otherwise."""
    P = recon_fit_trivis_full.ReconFitterTriVisFull
    fit = object.__new__(P)
    fit.outpath = str(tmp_path)
    imgs = [f"/data/Date03_x/t0001.000/k1.color.jpg", f"/data/Date03_x/t0002.000/k1.color.jpg"]
    assert not fit.is_done(imgs, "sn", 1) and not fit.is_done(imgs, "sn", 1, neural_only=True)
    for t in ("t0001.000", "t0002.000"):
        d = tmp_path / "Date03_x" / t / "sn"
        d.mkdir(parents=True)
        (d / "k1_densepc.npz").write_bytes(b"x")
    assert fit.is_done(imgs, "sn", 1, neural_only=True) and not fit.is_done(imgs, "sn", 1)
    for t in ("t0001.000", "t0002.000"):
        for n in ("k1.smpl.pkl", "k1.object.pkl"):
            (tmp_path / "Date03_x" / t / "sn" / n).write_bytes(b"x")
    assert fit.is_done(imgs, "sn", 1)
    f30 = object.__new__(fit_SMPLH_30fps.SMPLHFitter30fps)
    (tmp_path / "fr").mkdir()
    assert not f30.is_done(str(tmp_path / "fr"), 1)
    (tmp_path / "fr" / "k1.smplfit_temporal.pkl").write_bytes(b"0" * 50)
    assert not f30.is_done(str(tmp_path / "fr"), 1)                       # a stub of <= 100 bytes does not count (fit_SMPLH_kpts.py:340-345)
    (tmp_path / "fr" / "k1.smplfit_temporal.pkl").write_bytes(b"0" * 500)
    assert f30.is_done(str(tmp_path / "fr"), 1)
