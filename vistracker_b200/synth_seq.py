"""A seeded synthetic sequence written in the BEHAVE layout the reference's programs read (``shims/seqio.py`` lists the files): the raw inputs of
scripts/demo.sh -- colour frames, person / object masks, FrankMocap parameters, OpenPose key points, ``info.json`` -- for the integration tests
and ``tools/run_sequence.py``.  There is no network access for the real dataset; shapes, file names, value ranges and the coupling between
the files (masks that follow the projected body, key points that are projections of a joint cloud riding on the translation) are what the
programs rely on."""
from __future__ import annotations

import json
import os
from typing import Dict

import numpy as np

from .config import KINECT_CX_PX, KINECT_CY_PX, KINECT_FX_PX, KINECT_FY_PX
from .synth import synthetic_camera_frame
from .synth_smpl import synthetic_motion


def write_synthetic_sequence(root: str, frames: int = 24, seq_name: str = "Date03_Sub03_chairwood_synth", kid: int = 1, seed: int = 7,
                             H: int = 1536, W: int = 2048) -> Dict[str, object]:
    """Writes ``<root>/<seq_name>/`` and returns {'seq_folder', 'frames', 'pose', 'betas', 'trans', 'kpts'} (the motion behind the files)."""
    from PIL import Image
    rng = np.random.Generator(np.random.PCG64(seed))
    seq = os.path.join(root, seq_name)
    os.makedirs(seq, exist_ok=True)
    with open(os.path.join(seq, "info.json"), "w") as f:
        json.dump({"cat": "chairwood", "gender": "male", "config": None, "intrinsic": None, "empty": None, "beta": None, "kinects": [0, 1, 2, 3]}, f)
    pose, betas, trans = (t.numpy() for t in synthetic_motion(frames, seed=seed))
    off = rng.standard_normal((25, 3)) * np.array([0.25, 0.45, 0.12])
    names = [f"t{i / 30.0:08.3f}" for i in range(frames)]                       # 30 fps time stamps: t0000.000, t0000.033, ...
    kpts_all = []
    for i, name in enumerate(names):
        folder = os.path.join(seq, name)
        os.makedirs(folder, exist_ok=True)
        # the body centre projected into the image: masks and colour frame are placed around it
        cx = trans[i, 0] * KINECT_FX_PX / trans[i, 2] + KINECT_CX_PX
        cy = trans[i, 1] * KINECT_FY_PX / trans[i, 2] + KINECT_CY_PX
        rgb, person, obj = synthetic_camera_frame(H, W, seed=seed * 1000 + i % 4, center=(float(cx), float(cy)))
        Image.fromarray(rgb, "RGB").save(os.path.join(folder, f"k{kid}.color.jpg"), quality=92)
        Image.fromarray(person, "L").save(os.path.join(folder, f"k{kid}.person_mask.png"))
        Image.fromarray(obj, "L").save(os.path.join(folder, f"k{kid}.obj_rend_mask.png"))
        p72 = np.concatenate([pose[i, :66] + rng.standard_normal(66) * 0.05, np.zeros(6)])          # FrankMocap: 72-d SMPL pose, noisy
        with open(os.path.join(folder, f"k{kid}.mocap.json"), "w") as f:
            json.dump({"pose": p72.tolist(), "betas": (rng.standard_normal(10) * 0.3).tolist()}, f)
        J = trans[i][None] + off
        k2d = np.stack([J[:, 0] * KINECT_FX_PX / J[:, 2] + KINECT_CX_PX, J[:, 1] * KINECT_FY_PX / J[:, 2] + KINECT_CY_PX], -1) + rng.standard_normal((25, 2)) * 2.0
        conf = rng.uniform(0.4, 1.0, (25, 1)); conf[rng.random((25, 1)) < 0.1] = 0.05
        k = np.concatenate([k2d, conf], -1)
        kpts_all.append(k)
        with open(os.path.join(folder, f"k{kid}.color.json"), "w") as f:
            json.dump({"body_joints": k.reshape(-1).tolist()}, f)
    return {"seq_folder": seq, "frames": names, "pose": pose, "betas": betas, "trans": trans, "kpts": np.stack(kpts_all, 0)}
