"""``preprocess/fit_SMPLH_30fps.py`` -- demo.sh step 1: SMPL-T pre-fit of a sequence to OpenPose key points with temporal smoothness.
Same command line (preprocess/fit_SMPLH_30fps.py:215-233), ``main(args)`` and ``SMPLHFitter30fps.fit_seq(seq_folder, kid, start, end, redo,
bs)`` (preprocess/fit_SMPLH_kpts.py:84-112); the optimisation is ``vistracker_b200.fit_smplt`` (one CUDA-graph replay per Adam step)."""
from __future__ import annotations

import os
import os.path as osp
import sys
from argparse import ArgumentParser

import numpy as np
import torch

from .. import io as vio
from ..config import KINECT_CX_PX, KINECT_CY_PX, KINECT_FX_PX, KINECT_FY_PX
from ..fit_smplt import SMPLHFitter30fps as _Fitter30
from ..recon_fit import smplh_pose
from .assets import get_asset_provider
from .seqio import FrameDataReader


def get_parser() -> ArgumentParser:
    """preprocess/fit_SMPLH_30fps.py:218-227."""
    parser = ArgumentParser()
    parser.add_argument('-s', '--seq_folder')
    parser.add_argument('-d', '--debug', default=False, action='store_true')
    parser.add_argument('-fs', '--start', type=int, default=0)
    parser.add_argument('-fe', '--end', type=int, default=None)
    parser.add_argument('-redo', default=False, action='store_true')
    parser.add_argument('-i', '--init_type', choices=['mocap', 'pare'], default='mocap', help='source of init SMPL pose')
    parser.add_argument('-k', '--kid', default=1, type=int)
    parser.add_argument('-icap', default=False, action='store_true', help='If True, process InterCap dataset')
    parser.add_argument('-bs', '--batch_size', default=512, type=int)
    return parser


class SMPLHFitter30fps:
    SUFFIX = "smplfit_temporal"               # get_outfile (fit_SMPLH_30fps.py:202-203)
    SAVE_MESH = False                         # "not saving mesh for this" (fit_SMPLH_30fps.py:70-71)
    FITTER = _Fitter30
    smpl_depth = 2.2

    def __init__(self, debug=False, init_type='mocap', args=None, device="cuda:0"):
        self.debug, self.init_type, self.args = debug, init_type, args
        self.icap = bool(getattr(args, "icap", False))
        self.device = torch.device(device)
        self.fx, self.fy, self.cx, self.cy = KINECT_FX_PX, KINECT_FY_PX, KINECT_CX_PX, KINECT_CY_PX
        self.assets = get_asset_provider()
        # None = the reference's caps; VT_SHIM_SMPLT_MAX_ITER shortens the loop (integration tests)
        self.max_iter = int(os.environ["VT_SHIM_SMPLT_MAX_ITER"]) if os.environ.get("VT_SHIM_SMPLT_MAX_ITER") else None
        self._fitters = {}

    # ---- file conventions
    def get_outfile(self, frame_folder, kid):
        return osp.join(frame_folder, f'k{kid}.{self.SUFFIX}.pkl')

    def is_done(self, frame_folder, kid):
        """BaseFitter.is_done (fit_SMPLH_kpts.py:340-345): the result exists and is not a stub."""
        f = self.get_outfile(frame_folder, kid)
        return osp.isfile(f) and osp.getsize(f) > 100

    def is_batch_done(self, start, batch_end, reader, kid, redo):
        """fit_SMPLH_30fps.py:73-88."""
        if redo:
            return False
        return all(self.is_done(reader.get_frame_folder(i), kid) for i in range(start, batch_end))

    # ---- the reference's driver
    def fit_seq(self, seq_folder, kid, start, end, redo, bs=512):
        """fit_SMPLH_kpts.py:84-112: mini-batches of ``bs`` frames."""
        seq_name = osp.basename(seq_folder.rstrip('/'))
        if not self.icap:
            assert 'Date0' in seq_name or seq_name.startswith('S0'), "camera parameters are the BEHAVE / NTU-RGBD ones"
        reader = FrameDataReader(seq_folder)
        batch_end = reader.cvt_end(end)
        print(f"In total {(batch_end - start) // bs + 1} mini-batches.")
        if batch_end - start > bs:
            for bstart in range(start, batch_end, bs):
                self.fit_one_batch(seq_folder, kid, bstart, min(batch_end, bstart + bs), redo)
        else:
            self.fit_one_batch(seq_folder, kid, start, end, redo)

    def init_smpl(self, reader, kid, start, batch_end, redo):
        """fit_SMPLH_30fps.py:90-151: FrankMocap pose, fixed shape (beta_0 = 2.2), translation from the person-mask box at 2.2 m."""
        poses, trans, frame_inds = [], [], []
        for idx in range(start, batch_end):
            if self.is_done(reader.get_frame_folder(idx), kid) and not redo:
                continue
            p, _ = reader.get_mocap_params(idx, kid)
            mask = reader.get_mask(idx, kid, 'person')
            ys, xs = np.where(mask)
            if len(xs) < 10:
                raise ValueError(f"no person mask in {reader.get_frame_folder(idx)} kinect {kid}")
            bx = ((xs.max() + xs.min()) // 2 - self.cx) / self.fx * self.smpl_depth
            by = ((ys.max() + ys.min()) // 2 - self.cy) / self.fy * self.smpl_depth
            poses.append(p); trans.append(np.array([bx, by, self.smpl_depth])); frame_inds.append(idx)
        if not poses:
            return None, None
        betas = np.zeros((len(poses), 10), np.float32); betas[:, 0] = 2.2
        hand_mean = self.assets.priors(self.device).hand_mean
        return (smplh_pose(np.stack(poses, 0), hand_mean), torch.from_numpy(betas), torch.from_numpy(np.stack(trans, 0)).float()), frame_inds

    def load_kpts(self, reader, kid, frame_inds, tol=0.1):
        """fit_SMPLH_kpts.py:312-338."""
        kpts, files = [], []
        for idx in frame_inds:
            k = reader.get_body_kpts(idx, kid, tol)
            assert k is not None, f'{reader.get_frame_folder(idx)}/kinect {kid}'
            kpts.append(k); files.append(osp.join(reader.get_frame_folder(idx), f'k{kid}.color.jpg'))
        return torch.from_numpy(np.stack(kpts, 0)).float().to(self.device), files

    def fit_one_batch(self, seq_folder, kid, start, end, redo):
        reader = FrameDataReader(seq_folder)
        batch_end = reader.cvt_end(end)
        if self.is_batch_done(start, batch_end, reader, kid, redo):
            print(kid, 'all done')
            return
        init, frame_inds = self.init_smpl(reader, kid, start, batch_end, redo)
        if init is None:
            print(kid, 'all done')
            return
        kpts, image_files = self.load_kpts(reader, kid, frame_inds)
        pose0, betas0, trans0 = init
        assert len(kpts) == betas0.shape[0], f'kpts shape: {kpts.shape}, smpl betas shape: {betas0.shape}'
        print(f"Run SMPL-T fitting for {image_files[0]} -> {image_files[-1]}, batch size={len(kpts)}")
        gender = reader.seq_info.get_gender()
        if gender not in self._fitters:
            self._fitters[gender] = self.FITTER(self.assets.smplh(gender, self.device), self.assets.body25(self.device), self.assets.prior_arrays(), icap=self.icap)
        fitter = self._fitters[gender]
        fitter._graph = None if getattr(fitter, "_graph_B", None) != len(kpts) else fitter._graph     # one captured graph per batch size
        fitter._graph_B = len(kpts)
        res = fitter.fit_batch(pose0, betas0, trans0, kpts, max_iter=self.max_iter)
        self.save_results(res, gender, image_files, kid)
        return res

    def save_results(self, res, gender, image_files, kid):
        """fit_SMPLH_kpts.py:228-261 (``skip_frame`` keeps every frame here, fit_SMPLH_30fps.py:67-68)."""
        outfiles = [self.get_outfile(osp.dirname(f), kid) for f in image_files]
        vio.save_smplt_fits(outfiles, res["pose"], res["betas"], res["trans"])
        if self.SAVE_MESH:
            layer = self.assets.smplh(gender, self.device)
            with torch.no_grad():
                verts = layer(res["pose"], th_betas=res["betas"], th_trans=res["trans"])[0]
            for f, v in zip(outfiles, verts):
                vio.save_ply(f.replace('.pkl', '.ply'), v, layer.faces)


def main(args):
    fitter = SMPLHFitter30fps(debug=args.debug, init_type=args.init_type, args=args)
    fitter.fit_seq(args.seq_folder, args.kid, args.start, args.end, args.redo, args.batch_size)
    print("all done")


def cli(argv=None) -> int:
    """The reference swallows every exception and exits 0 (fit_SMPLH_30fps.py:229-233); here a failure is printed AND reported."""
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
