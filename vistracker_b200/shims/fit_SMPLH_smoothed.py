"""``preprocess/fit_SMPLH_smoothed.py`` -- demo.sh step 2b: re-fit starting from the SmoothNet output pack
(``RECON_PATH/recon_<smoothed_name>/<seq>_k<kid>.pkl``), writes ``k<kid>.smplfit_smoothed.pkl`` + ``.ply`` (the mesh step 3 renders)."""
from __future__ import annotations

import os.path as osp
import sys
from argparse import ArgumentParser

import numpy as np
import torch

from .. import io as vio
from ..fit_smplt import SMPLHFitterSmoothed as _FitterSm
from ..recon_fit import smplh_pose
from . import paths
from .fit_SMPLH_30fps import SMPLHFitter30fps
from .seqio import FrameDataReader


def get_parser() -> ArgumentParser:
    """preprocess/fit_SMPLH_smoothed.py:126-137."""
    parser = ArgumentParser()
    parser.add_argument('-s', '--seq_folder')
    parser.add_argument('-d', '--debug', default=False, action='store_true')
    parser.add_argument('-fs', '--start', type=int, default=0)
    parser.add_argument('-fe', '--end', type=int, default=None)
    parser.add_argument('-redo', default=False, action='store_true')
    parser.add_argument('-i', '--init_type', default='mocap', choices=['mocap', 'pare'])
    parser.add_argument('-k', '--kid', default=1, type=int)
    parser.add_argument('-sn', '--smoothed_name', default='smplt-smoothed', help='save name of the SmoothNet pack to start from')
    parser.add_argument('-icap', default=False, action='store_true')
    parser.add_argument('-bs', '--batch_size', default=512, type=int)
    return parser


class SMPLHFitterSmoothed(SMPLHFitter30fps):
    SUFFIX = "smplfit_smoothed"
    SAVE_MESH = True                          # save_smpl_mesh writes the ply (fit_SMPLH_smoothed.py:68-69)
    FITTER = _FitterSm

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        p = paths()
        self.packed_path, self.gtpack_path = p["RECON_PATH"], p.get("GT_PACKED")

    def init_smpl(self, reader, kid, start, batch_end, redo):
        """fit_SMPLH_smoothed.py:26-66: every frame of the mini-batch from the pack (no per-frame skipping)."""
        name = self.args.smoothed_name
        pack = vio.load_packed(osp.join(self.packed_path, f'recon_{name}/{reader.seq_name}_k{kid}.pkl'))
        if list(pack["frames"]) != reader.frames:
            raise ValueError(f"the frames of recon_{name} do not match {reader.seq_path}")
        hand_mean = self.assets.priors(self.device).hand_mean
        sl = slice(start, batch_end)
        f = lambda a: torch.from_numpy(np.asarray(a[sl], np.float32))
        return (smplh_pose(np.asarray(pack["poses"][sl]), hand_mean), f(pack["betas"]), f(pack["trans"])), list(range(start, batch_end))

    def load_kpts(self, reader, kid, frame_inds, tol=0.1):
        """fit_SMPLH_smoothed.py:84-110: ``joints2d`` of the GT pack when there is one, else the per-frame OpenPose files."""
        f = osp.join(self.gtpack_path, f'{reader.seq_name}_GT-packed.pkl') if self.gtpack_path else None
        if not f or not osp.isfile(f):
            print(f"Warning: no packed GT data found in {f}! Loading separate J2d data.")
            return super().load_kpts(reader, kid, frame_inds, tol)
        pack = vio.load_packed(f)
        if list(pack["frames"]) != reader.frames:
            raise ValueError(f"the frames of {f} do not match {reader.seq_path}")
        kpts = np.asarray(pack['joints2d'])[frame_inds, kid].astype(np.float32).copy()
        kpts[:, :, 2][kpts[:, :, 2] < tol] = 0
        files = [osp.join(reader.get_frame_folder(i), f'k{kid}.color.jpg') for i in frame_inds]
        return torch.from_numpy(kpts).to(self.device), files


def main(args):
    fitter = SMPLHFitterSmoothed(debug=args.debug, init_type=args.init_type, args=args)
    fitter.fit_seq(args.seq_folder, args.kid, args.start, args.end, args.redo, args.batch_size)
    print("all done")


def cli(argv=None) -> int:
    import traceback
    args = get_parser().parse_args(argv)
    try:
        main(args)
    except Exception:
        print(traceback.format_exc())
        return 1
    return 0


if __name__ == '__main__':
    sys.exit(cli())
