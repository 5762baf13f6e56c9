"""demo.sh steps 1, 2b, 3, 4 and 6 through the entry-point shims on a synthetic BEHAVE-layout sequence folder (vistracker_b200/synth_seq.py),
with the stages in between (SmoothNet, the packs) done by library calls as tools/run_sequence.py does: every program finds the files the
previous one wrote, writes the files the next one reads, skips finished work on a second run and honours -redo.  Loops are shortened."""
import json
import os
import pickle as pkl

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_demo_steps_through_the_shims(tmp_path, monkeypatch, capsys):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import joblib
    from vistracker_b200 import io as vio
    from vistracker_b200.shims import assets, fit_SMPLH_30fps, fit_SMPLH_smoothed, recon_fit_trivis_full, render_triplane_nr
    from vistracker_b200.synth_seq import write_synthetic_sequence
    T = 20
    d = write_synthetic_sequence(str(tmp_path / "data"), frames=T)
    seq, frames = d["seq_folder"], d["frames"]
    seq_name = os.path.basename(seq)
    recon = str(tmp_path / "recon")
    monkeypatch.setenv("VT_RECON_PATH", recon)
    monkeypatch.setenv("VT_SHIM_SMPLT_MAX_ITER", "3")
    monkeypatch.setenv("VT_SHIM_RECON_LOOP", json.dumps({"max_iter": 1, "steps_per_iter": 2}))
    monkeypatch.chdir(tmp_path)                                               # no PATHS.yml / config folder here: built-in options
    assets.set_asset_provider(assets.SyntheticAssets(os.path.join(HERE, "golden", "assets.npz")))
    try:
        # ---- step 1: python preprocess/fit_SMPLH_30fps.py -s ${seq} -bs 512
        assert fit_SMPLH_30fps.cli(["-s", seq, "-bs", "512"]) == 0
        fits = [os.path.join(seq, f, "k1.smplfit_temporal.pkl") for f in frames]
        assert all(os.path.getsize(f) > 100 for f in fits)
        p0 = pkl.load(open(fits[3], "rb"))
        assert p0["pose"].shape == (156,) and p0["betas"].shape == (10,) and p0["trans"].shape == (3,) and abs(p0["trans"][2] - 2.2) < 0.5
        stamps = [os.path.getmtime(f) for f in fits]
        assert fit_SMPLH_30fps.cli(["-s", seq, "-bs", "512"]) == 0            # second run: every frame is done
        assert [os.path.getmtime(f) for f in fits] == stamps and "all done" in capsys.readouterr().out
        assert fit_SMPLH_30fps.cli(["-s", seq, "-bs", "8", "-redo", "-fe", "12"]) == 0        # -redo, two mini-batches of 8 and 4 frames
        after = [os.path.getmtime(f) for f in fits]
        assert all(a > b for a, b in zip(after[:12], stamps[:12])) and after[12:] == stamps[12:]
        # ---- step 2a (stand-in for smoothnet/smooth_smplt.py, which needs >= 64 frames and is tested in test_gpu_smooth.py): a moving average
        #      of the step-1 trajectory in the pack format SmoothNet writes -> recon_smplt-smoothed
        poses, betas, trans = vio.load_smplt_fits(fits)
        k = np.ones(5) / 5
        avg = lambda a: np.stack([np.convolve(np.pad(a[:, j], 2, mode="edge"), k, mode="valid") for j in range(a.shape[1])], 1).astype(np.float32)
        vio.pack_smplt(os.path.join(recon, "recon_smplt-smoothed", f"{seq_name}_k1.pkl"), frames, "male", avg(poses), betas, avg(trans))
        # ---- step 2b: python preprocess/fit_SMPLH_smoothed.py -sn smplt-smoothed -s ${seq}
        assert fit_SMPLH_smoothed.cli(["-sn", "smplt-smoothed", "-s", seq]) == 0
        sfits = [os.path.join(seq, f, "k1.smplfit_smoothed.pkl") for f in frames]
        assert all(os.path.isfile(f) and os.path.isfile(f.replace(".pkl", ".ply")) for f in sfits)
        v, fc = vio.load_ply(sfits[0].replace(".pkl", ".ply"))
        assert v.shape == (6890, 3) and fc.shape[1] == 3
        # ---- step 2c (library): python preprocess/pack_smplt.py -t 1 -m smoothed
        poses, betas, trans = vio.load_smplt_fits(sfits)
        vio.pack_smplt(os.path.join(recon, "recon_smplt-smoothed-fit", f"{seq_name}_k1.pkl"), frames, "male", poses, betas, trans)
        # ---- step 3: python render/render_triplane_nr.py -s ${seq}
        assert render_triplane_nr.cli(["-s", seq]) == 0
        tri = vio.load_triplane_png(os.path.join(seq, frames[5], "k1.smooth_triplane.png"))
        assert tri.shape == (512, 512, 3) and set(np.unique(tri)) <= {0, 255} and tri.any()
        # ---- step 4: python recon/recon_fit_trivis_full.py tri-vis-l2 -sn test-release -or neural -sr smplt-smoothed-fit -t 1 -bs 64 -tt smooth -neural_only
        #      (-fv 10: the random-init UDF never falls under the production 0.004)
        step4 = ["tri-vis-l2", "-sn", "test-release", "-or", "neural", "-sr", "smplt-smoothed-fit", "-t", "1", "-bs", "64", "-tt", "smooth", "-neural_only",
                 "-fv", "10", "-s", seq]
        assert recon_fit_trivis_full.cli(step4) == 0
        npz = os.path.join(recon, seq_name, frames[7], "test-release", "k1_densepc.npz")
        dense = np.load(npz, allow_pickle=True)
        obj = dense["object"].item()
        assert obj["points"].shape[1] == 3 and obj["pca_axis"].shape == (3, 3) and obj["centers"].shape == (6,) and np.isnan(obj["centers"][:3]).all()
        capsys.readouterr()
        assert recon_fit_trivis_full.cli(step4) == 0 and "already done, skipped" in capsys.readouterr().out
        # ---- step 5 (stand-in for SmoothNet + HVOP-Net): an object-rotation pack under the name step 6 loads
        R = torch.linalg.qr(torch.randn(T, 3, 3, generator=torch.Generator().manual_seed(3)))[0]
        R = R * torch.sign(torch.linalg.det(R))[:, None, None]
        os.makedirs(os.path.join(recon, "recon_smooth-hvopnet"), exist_ok=True)
        joblib.dump({"frames": frames, "obj_angles": R.numpy(), "obj_trans": np.zeros((T, 3)), "obj_scales": np.ones(T)},
                    os.path.join(recon, "recon_smooth-hvopnet", f"{seq_name}_k1.pkl"))
        # ---- step 6: python recon/recon_fit_trivis_full.py tri-vis-l2 -sr smplt-smoothed-fit -or smooth-hvopnet -sn test-releasev2 -s ${seq}
        step6 = ["tri-vis-l2", "-sr", "smplt-smoothed-fit", "-or", "smooth-hvopnet", "-sn", "test-releasev2", "-fv", "10", "-bs", "12", "-s", seq]
        assert recon_fit_trivis_full.cli(step6) == 0
        for f in (frames[0], frames[11], frames[12], frames[-1]):             # two batches: 12 + 8 frames
            folder = os.path.join(recon, seq_name, f, "test-releasev2")
            s = pkl.load(open(os.path.join(folder, "k1.smpl.pkl"), "rb"))
            o = pkl.load(open(os.path.join(folder, "k1.object.pkl"), "rb"))
            assert s["pose"].shape == (156,) and s["betas"].shape == (10,) and s["trans"].shape == (3,)
            assert np.abs(o["rot"] @ o["rot"].T - np.eye(3)).max() < 1e-5 and abs(np.linalg.det(o["rot"]) - 1) < 1e-5 and o["trans"].shape == (3,)
        capsys.readouterr()
        assert recon_fit_trivis_full.cli(step6) == 0 and capsys.readouterr().out.count("already done, skipped") == 2
        # a failing run reports it to the shell (the reference prints the traceback and exits 0)
        assert recon_fit_trivis_full.cli(["tri-vis-l2", "-sr", "no-such-pack", "-or", "neural", "-sn", "x", "-fv", "10", "-s", seq]) == 1
    finally:
        assets.set_asset_provider(None)
