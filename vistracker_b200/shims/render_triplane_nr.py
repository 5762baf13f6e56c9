"""``render/render_triplane_nr.py`` -- demo.sh step 3: the smoothed SMPL-T mesh of every frame rendered as three orthographic silhouettes
(right, back, top) into ``k<kid>.smooth_triplane.png`` (render_triplane_nr.py:37-107)."""
from __future__ import annotations

import os.path as osp
import sys
from argparse import ArgumentParser

import numpy as np
import torch

from .. import io as vio
from ..render import TriplaneNrRenderer as _Renderer
from .assets import get_asset_provider
from .seqio import FrameDataReader


def get_parser() -> ArgumentParser:
    """render/render_triplane_nr.py:153-161."""
    parser = ArgumentParser()
    parser.add_argument('-s', "--seq_folder")
    parser.add_argument('-fs', '--start', type=int, default=0)
    parser.add_argument('-fe', '--end', type=int, default=None)
    parser.add_argument('-k', '--kids', default=[1], nargs='+', type=int)
    parser.add_argument('-redo', default=False, action='store_true')
    parser.add_argument('-sn', '--smpl_name', help='smpl fitting save name', default='fit02')
    parser.add_argument('-on', '--obj_name', help='object fitting save name', default='fit01')
    parser.add_argument('-mesh_type', default='smooth')
    return parser


# mesh / image names per mesh type (render_triplane_nr.py:60-83, data/testdata_triplane.py:84-104)
_FILES = {"smooth": ("smplfit_smoothed.ply", "smooth_triplane.png"), "temporal": ("smplfit_temporal.ply", "mocap_triplane.png"),
          "mocap": ("smplfit_kpt.ply", "mocap_triplane.png")}


class TriplaneNrRenderer:
    def __init__(self, image_size=512, device="cuda:0", batch=64):
        self.device = torch.device(device)
        self.r = _Renderer(image_size, self.device)
        self.body25 = get_asset_provider().body25(self.device)
        self.batch = batch

    def render_seq(self, seq_folder, start, end, kids, mesh_type='smooth', smpl_name='fit02', obj_name='fit01', redo=False):
        if mesh_type not in _FILES:
            raise ValueError(f"mesh type {mesh_type}: the ground-truth / fit02 variants need the BEHAVE registrations")
        mesh_ext, img_ext = _FILES[mesh_type]
        reader = FrameDataReader(seq_folder)
        batch_end = reader.cvt_end(end)
        for kid in kids:
            todo = []
            for idx in range(start, batch_end):
                out = osp.join(reader.get_frame_folder(idx), f'k{kid}.{img_ext}')
                if osp.isfile(out) and not redo:                     # "already exists, skipped" (render_triplane_nr.py:52-54)
                    continue
                todo.append((osp.join(reader.get_frame_folder(idx), f'k{kid}.{mesh_ext}'), out))
            for s in range(0, len(todo), self.batch):
                chunk = todo[s:s + self.batch]
                meshes = [vio.load_ply(m) for m, _ in chunk]
                faces = meshes[0][1]
                verts = torch.from_numpy(np.stack([m[0] for m in meshes], 0)).float().to(self.device)
                center = self.body25(verts)[:, 8]                        # get_smpl_center: body-25 joint 8
                masks = self.r.render_3views(faces, verts - center[:, None])
                vio.save_triplane_png([o for _, o in chunk], masks)
        print('all done')


def main(args):
    tri_renderer = TriplaneNrRenderer()
    tri_renderer.render_seq(args.seq_folder, args.start, args.end, args.kids, args.mesh_type, args.smpl_name, args.obj_name, redo=args.redo)


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
