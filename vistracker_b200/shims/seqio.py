"""The BEHAVE sequence layout as the reference's programs see it, file names only -- the slice of the reference's vendored ``behave``
toolkit (behave/frame_data.py, sync_frame.py, seq_utils.py) and of ``data/base_data.py`` that demo.sh steps 1-6 touch:

    SEQ/info.json                       {"cat", "gender", "config", "intrinsic", "empty", "beta", "kinects"}      (seq_utils.py:51-64)
    SEQ/t0003.000/                      one folder per frame, sorted                                               (sync_frame.py:29-54)
        k1.color.jpg                    RGB frame 2048 x 1536                                                      (sync_frame.py:61)
        k1.person_mask.png|jpg          k1.obj_rend_mask.png|jpg  (or k1.obj_mask.*)                               (frame_data.py:198-215, base_data.py:96-131)
        k1.mocap.json                   FrankMocap {"pose": [72], "betas": [10]}                                   (frame_data.py:92-97)
        k1.color.json                   OpenPose {"body_joints": [25 * 3]}                                         (frame_data.py:172-183)
        k1.smplfit_temporal.pkl         step 1 output; k1.smplfit_smoothed.pkl / .ply step 2; k1.smooth_triplane.png step 3
    RECON_PATH/recon_<name>/<seq>_k<kid>.pkl   joblib packs of pack_smplt.py / pack_recon.py / SmoothNet / HVOP-Net
    RECON_PATH/<seq>/<frame>/<save_name>/k<kid>_densepc.npz | k<kid>.smpl.pkl | k<kid>.object.pkl                  (recon_fit_base.py:260-313)

Decoding is Pillow's (the reference mixes PIL and cv2); everything after decoding runs on the device.
"""
from __future__ import annotations

import json
import os
from os.path import basename, isdir, isfile, join
from typing import List, Optional, Sequence

import numpy as np


class SeqInfo:
    """behave/seq_utils.py:11-64."""

    def __init__(self, seq_path: str):
        with open(join(seq_path, "info.json")) as f:
            self.info = json.load(f)

    def get_obj_name(self):
        return self.info["cat"]

    def get_gender(self):
        return self.info["gender"]

    @property
    def kids(self):
        return self.info.get("kinects", [0, 1, 2, 3])


class FrameDataReader:
    """behave/frame_data.py: the accessors the fitters call (frames, folders, masks, FrankMocap parameters, OpenPose key points)."""

    def __init__(self, seq: str, check_image: bool = False):
        self.seq_path = seq.rstrip(os.sep)
        self.seq_name = basename(self.seq_path)
        self.seq_info = SeqInfo(self.seq_path)
        self.kids = self.seq_info.kids
        # sync_frame.py:29-54: every sub-folder whose name starts with 't' (a time stamp), sorted
        self.frames = sorted(d for d in os.listdir(self.seq_path) if d.startswith("t") and isdir(join(self.seq_path, d)))
        if check_image:
            self.frames = [f for f in self.frames if isfile(join(self.seq_path, f, f"k{self.kids[0]}.color.jpg"))]

    def __len__(self):
        return len(self.frames)

    def cvt_end(self, end: Optional[int]) -> int:
        """frame_data.py:238-242."""
        return len(self) if end is None or end > len(self) else end

    def get_frame_folder(self, idx) -> str:
        return join(self.seq_path, self.frames[idx] if isinstance(idx, int) else idx)

    def get_color_files(self, idx, kids: Sequence[int]) -> List[str]:
        return [join(self.get_frame_folder(idx), f"k{k}.color.jpg") for k in kids]

    def get_mask_file(self, idx, kid: int, cat: str) -> str:
        """frame_data.py:198-215."""
        folder = self.get_frame_folder(idx)
        if cat == "person":
            f = join(folder, f"k{kid}.person_mask.png")
            return f if isfile(f) else join(folder, f"k{kid}.person_mask.jpg")
        if cat != "obj":
            raise NotImplementedError(cat)
        f = ""
        for ext in ("png", "jpg"):
            f = join(folder, f"k{kid}.obj_rend_mask.{ext}")
            if not isfile(f):
                f = join(folder, f"k{kid}.obj_mask.{ext}")
            if isfile(f):
                break
        return f

    def get_mask(self, idx, kid: int, cat: str = "person", ret_bool: bool = True):
        f = self.get_mask_file(idx, kid, cat)
        if not isfile(f):
            return None
        m = read_gray(f)
        return m > 127 if ret_bool else m

    def get_mocap_params(self, idx, kid: int = 1):
        f = join(self.get_frame_folder(idx), f"k{kid}.mocap.json")
        if not isfile(f):
            return None, None
        with open(f) as fh:
            p = json.load(fh)
        return np.array(p["pose"]), np.array(p["betas"])

    def get_body_kpts(self, idx, kid: int, tol: float = 0.5):
        f = join(self.get_frame_folder(idx), f"k{kid}.color.json")
        if not isfile(f):
            return None
        with open(f) as fh:
            J2d = np.array(json.load(fh)["body_joints"], np.float64).reshape(-1, 3)
        J2d[:, 2][J2d[:, 2] < tol] = 0
        return J2d


def read_rgb(file: str) -> np.ndarray:
    """[H, W, 3] uint8 RGB (``BaseDataset.load_rgb``, data/base_data.py:173-183: ``cv2.imread(rgb_file)[:, :, ::-1]``)."""
    from PIL import Image
    return np.asarray(Image.open(file).convert("RGB"))


def read_gray(file: str) -> np.ndarray:
    """[H, W] uint8 (``np.array(Image.open(mask_file))`` / ``cv2.imread(file, cv2.IMREAD_GRAYSCALE)``)."""
    from PIL import Image
    return np.asarray(Image.open(file).convert("L"))


def load_masks(rgb_file: str):
    """``BaseDataset.load_masks`` (data/base_data.py:96-131): person mask png else jpg; first existing of the four object-mask names."""
    person = rgb_file.replace(".color.jpg", ".person_mask.png")
    if not isfile(person):
        person = rgb_file.replace(".color.jpg", ".person_mask.jpg")
    obj = None
    for pat in (".obj_rend_mask.png", ".obj_rend_mask.jpg", ".obj_mask.png", ".obj_mask.jpg"):
        obj = rgb_file.replace(".color.jpg", pat)
        if isfile(obj):
            break
    return read_gray(person), read_gray(obj)


def load_kpts_json(json_paths: Sequence[str], tol: float) -> np.ndarray:
    """``ReconFitterBase.load_kpts`` (recon/recon_fit_base.py:381-396): [B, 25, 3] OpenPose body joints in original image pixels."""
    out = []
    for f in json_paths:
        with open(f) as fh:
            J2d = np.array(json.load(fh)["body_joints"], np.float64).reshape(-1, 3)
        J2d[:, 2][J2d[:, 2] < tol] = 0
        out.append(J2d)
    return np.stack(out, 0).astype(np.float32)


def check_frame_consistency(packed: dict, seq_folder: str) -> bool:
    """``check_frame_consistency`` of the fitters: the pack's frame list is the folder's."""
    return list(packed["frames"]) == FrameDataReader(seq_folder).frames
