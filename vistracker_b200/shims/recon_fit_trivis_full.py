"""``recon/recon_fit_trivis_full.py`` -- demo.sh steps 4 (``-neural_only``) and 6: SIF-Net neural reconstruction and the joint SMPL-T / object
optimisation of a sequence, batch by batch.  Same command line (``get_parser``: recon_fit_triplane.py:242-266 + recon_fit_trivis_full.py:460-468),
``merge_configs`` (recon_fit_triplane.py:268-300, recon_fit_trivis_full.py:470-475), ``recon_fit(args)`` (:477-485) and ``fit_recon`` with its
skip-if-done / ``-redo`` logic (recon_fit_triplane.py:30-111, recon_fit_base.py:260-276); the per-batch work is
``vistracker_b200.recon_driver.fit_recon_batch``.  Frames are decoded on the host (Pillow); crop, resize, compositing, the network and both
optimisation loops run on the device."""
from __future__ import annotations

import json
import os
import os.path as osp
import sys
import time
from argparse import ArgumentParser

import numpy as np
import torch

from .. import io as vio
from ..config import default_options, load_configs as _load_configs
from ..frameio import prepare_images
from ..generator import GeneratorTriplaneVis
from ..geom import compute_pca
from ..recon_driver import fit_recon_batch, scale_body_kpts
from ..recon_fit import ReconFitterTriVisFull as _CoreFitter, SMPLParams, smplh_pose
from . import paths
from .assets import get_asset_provider
from .seqio import FrameDataReader, SeqInfo, load_kpts_json, load_masks, read_rgb


def load_configs(exp_name: str):
    """``config.config_loader.load_configs`` (config/config_loader.py:24-45): ``config/<exp_name>.json`` relative to the working directory when
    it is there (a checkout of the reference), else the built-in tri-vis-l2 options."""
    if osp.isfile(osp.join("config", exp_name + ".json")):
        return _load_configs(exp_name)
    if exp_name != "tri-vis-l2":
        raise FileNotFoundError(f"config/{exp_name}.json (only tri-vis-l2 is built in)")
    return default_options()


class ReconFitterTriVisFull:
    @staticmethod
    def get_parser():
        "return cmd line argument parser"
        parser = ArgumentParser()
        parser.add_argument('exp_name', help='experiment name')
        parser.add_argument('-s', '--seq_folder', help="path to one BEHAVE sequence")
        parser.add_argument('-sn', '--save_name', required=True, help='recon result save name')
        parser.add_argument('-o', '--outpath', default=paths()["RECON_PATH"], help='where to save reconstruction results')
        parser.add_argument('-ck', '--checkpoint', default=None, help='load which checkpoint, will find best or last checkpoint if None')
        parser.add_argument('-fv', '--filter_val', type=float, default=0.004, help='threshold value to filter surface points')
        parser.add_argument('-st', '--sparse_thres', type=float, default=0.03, help="filter value to filter sparse point clouds")
        parser.add_argument('-t', '--tid', default=1, type=int, help='test on images from which kinect')
        parser.add_argument('-bs', '--batch_size', default=96, type=int, help='optimization batch size')
        parser.add_argument('-redo', default=False, action='store_true')
        parser.add_argument('-d', '--display', default=False, action='store_true')
        parser.add_argument('-fs', '--start', default=0, type=int, help='start fitting from which frame')
        parser.add_argument('-fe', '--end', default=None, type=int, help='end fitting at which frame')
        parser.add_argument('-tt', '--triplane_type', default='smooth', choices=['gt', 'mocap', 'temporal', "smooth"],
                            help='use which triplane rendering results, for file names, see data/testdata_triplane.py')
        parser.add_argument('-pat', default='t*', help='pattern to get image files')
        parser.add_argument('-neural_only', default=False, action='store_true', help="Run SIF-Net neural prediction only")
        parser.add_argument('-pred_occ', default=True, action='store_true', help="use predicted occlusion ratio(visibility)")
        parser.add_argument('-sr', '--smpl_recon_name', required=True, help="SMPL-T result: used to initialize SMPL pose for joint opt")
        parser.add_argument('-or', '--obj_recon_name', required=True, help="Object pose used to initialize joint optimization")
        return parser

    @staticmethod
    def merge_configs(args, configs):
        """merge command line argument with network training configurations (recon_fit_triplane.py:268-300 + recon_fit_trivis_full.py:470-475)"""
        configs.batch_size = args.batch_size
        configs.test_kid = args.tid
        configs.filter_val = args.filter_val
        configs.sparse_thres = args.sparse_thres
        configs.seq_folder = args.seq_folder
        configs.pat = args.pat
        configs.save_name = args.save_name
        configs.checkpoint = args.checkpoint
        configs.outpath = args.outpath
        configs.redo = args.redo
        configs.display = args.display
        configs.start = args.start
        configs.end = args.end
        configs.neural_only = args.neural_only
        configs.pred_occ = args.pred_occ
        configs.triplane_type = args.triplane_type
        print("Triplane SMPL is from", args.triplane_type)
        configs.smpl_recon_name = args.smpl_recon_name
        configs.obj_recon_name = args.obj_recon_name
        return configs

    # file names per triplane type (data/testdata_triplane.py:84-104)
    _TRI = {"smooth": ("smooth_triplane.png", "smplfit_smoothed.ply"), "temporal": ("mocap_triplane.png", "smplfit_temporal.ply"),
            "mocap": ("mocap_triplane.png", "smplfit_kpt.ply")}

    def __init__(self, seq_folder, device='cuda:0', debug=False, obj_name=None, outpath=None, args=None):
        """recon_fit_base.py:54-110: object name and gender from ``info.json``, the template's PCA axes and 3000 surface samples, part labels."""
        self.args, self.seq_folder, self.debug = args, seq_folder.rstrip('/'), debug
        self.outpath = outpath if outpath is not None else paths()["RECON_PATH"]
        self.device = torch.device(device)
        self.assets = get_asset_provider()
        if osp.isfile(osp.join(self.seq_folder, 'info.json')):
            info = SeqInfo(self.seq_folder)
            obj_name, self.gender = info.get_obj_name(), info.get_gender()
        else:
            assert obj_name is not None, 'must provide the name of the object to be reconstructed!'
            self.gender = 'male'
        self.scan = self.assets.object_template(obj_name)
        self.pca_init = torch.from_numpy(compute_pca(self.scan[0])).float()
        self.obj_points = torch.from_numpy(self._sample_surface(self.scan[0], self.scan[1], 3000)).float()
        # empty = the reference's schedule; VT_SHIM_RECON_LOOP='{"max_iter": 2, ...}' shortens the loops (integration tests)
        self.loop_kw = json.loads(os.environ.get("VT_SHIM_RECON_LOOP") or "{}")
        self.neural_points = 4000

    @staticmethod
    def _sample_surface(v, f, n, seed=0):
        """``trimesh.Trimesh.sample(n)``: area-weighted faces, uniform barycentric coordinates (seeded here)."""
        rng = np.random.default_rng(seed)
        a, b, c = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
        area = np.linalg.norm(np.cross(b - a, c - a), axis=1)
        idx = rng.choice(len(f), size=n, p=area / area.sum())
        r1, r2 = np.sqrt(rng.random(n))[:, None], rng.random(n)[:, None]
        return ((1 - r1) * a[idx] + r1 * (1 - r2) * b[idx] + r1 * r2 * c[idx]).astype(np.float32)

    # ---- outputs (recon_fit_base.py:260-313)
    def is_done(self, image_paths, save_name, test_id, neural_only=False):
        for x in image_paths:
            parts = str(x).split(os.sep)
            folder = osp.join(self.outpath, parts[-3], parts[-2], save_name)
            files = [f'k{test_id}_densepc.npz'] if neural_only else [f'k{test_id}.smpl.pkl', f'k{test_id}.object.pkl']
            if not all(osp.isfile(osp.join(folder, f)) for f in files):
                return False
        return True

    def get_test_files(self, args):
        """recon_fit_base.py:411-419 / DataPaths.get_image_paths_seq: ``<seq>/<frame>/k<tid>.color.jpg`` of frames[start:end]."""
        reader = FrameDataReader(self.seq_folder)
        files = [osp.join(reader.get_frame_folder(i), f'k{args.test_kid}.color.jpg') for i in range(len(reader))]
        end = args.end if args.end is not None else len(files)
        return files[args.start:end]

    # ---- one batch of the data loader (data/testdata_triplane.py:42-74, data/train_data.py:143-162)
    def load_batch(self, files, triplane_type):
        img_ext, mesh_ext = self._TRI[triplane_type]
        dev = self.device
        rgb, person, obj, tri, verts = [], [], [], [], []
        for f in files:
            rgb.append(read_rgb(f))
            pm, om = load_masks(f)
            person.append(pm); obj.append(om)
            tri.append(vio.load_triplane_png(f.replace('color.jpg', img_ext)))
            mesh = f.replace('color.jpg', mesh_ext)
            if not osp.isfile(mesh):
                print(mesh, 'does not exist!')
                raise ValueError()
            verts.append(vio.load_ply(mesh)[0])
        up = lambda a: torch.from_numpy(np.ascontiguousarray(np.stack(a, 0))).to(dev)
        images, crop_center = prepare_images(up(rgb), up(person), up(obj), up(tri))
        body_center = self.assets.body25(dev)(up(verts).float())[:, 8].contiguous()
        return {"images": images, "crop_center": crop_center, "body_center": body_center, "path": list(files)}

    def fit_recon(self, args):
        """recon_fit_triplane.py:30-111."""
        dev = self.device
        files = self.get_test_files(args)
        print(f"In total {(len(files) + args.batch_size - 1) // args.batch_size} batches, {len(files)} images")
        net = self.assets.sifnet(args, dev)
        generator = GeneratorTriplaneVis(net, threshold=2.0, sparse_thres=args.sparse_thres, filter_val=args.filter_val)
        prior_arrays = self.assets.prior_arrays()
        fitter = _CoreFitter(net, self.assets.priors(dev), torch.from_numpy(prior_arrays["part_labels"].astype(np.int64)), scan=self.scan)
        layer, body25 = self.assets.smplh(self.gender, dev), self.assets.body25(dev)
        seq_name = osp.basename(self.seq_folder)
        packs = {}

        def pack(name):                                   # load_old_recon_packed (recon_fit_triplane.py:165-174)
            if name not in packs:
                packs[name] = vio.load_packed(osp.join(self.outpath, f'recon_{name}/{seq_name}_k1.pkl'))
            return packs[name]
        done = 0
        for i in range(0, len(files), args.batch_size):
            batch = files[i:i + args.batch_size]
            start_time = time.time()
            if self.is_done(batch, args.save_name, args.test_kid, args.neural_only) and not args.redo:
                print(f"{batch[0]}-{batch[-1]}", args.save_name, 'already done, skipped')
                continue
            data = self.load_batch(batch, args.triplane_type)
            folders = vio.output_folders(self.outpath, batch, args.save_name)

            def save_neural(s, e, pc):                    # save_neural_recon per mini-batch (recon_fit_behave.py:139-147)
                vio.save_neural_recon(folders[s:e], args.test_kid, pc)
            if args.neural_only:
                fit_recon_batch(fitter, generator, data, None, None, self.obj_points, neural_only=True, on_mini_batch=save_neural)
                print(f"Only saving neural reconstruction results for batch {i // args.batch_size}")
                done += len(batch)
                continue
            sm = vio.packed_batch(pack(args.smpl_recon_name), batch, args.test_kid)
            pose = smplh_pose(sm["poses"], fitter.priors.hand_mean)
            betas, trans = torch.from_numpy(sm["betas"]).float(), torch.from_numpy(sm["trans"]).float()
            init = lambda human_t: SMPLParams(layer, body25, pose.to(dev), betas.to(dev), trans.to(dev))      # get_smpl_init (recon_fit_trivis_full.py:62-77)
            kpts = torch.from_numpy(load_kpts_json([f.replace('.color.jpg', '.color.json') for f in batch], 0.3)).to(dev)
            body_kpts = scale_body_kpts(kpts, data["crop_center"])
            rot_init = None
            if args.obj_recon_name != 'neural':                 # HVOP-Net or other results (recon_fit_trivis_full.py:94-98)
                print(f'object rotation is from {args.obj_recon_name}')
                rot_init = torch.from_numpy(vio.packed_batch(pack(args.obj_recon_name), batch, args.test_kid)["obj_angles"]).float()
            out = fit_recon_batch(fitter, generator, data, init, body_kpts, self.obj_points, pca_init=self.pca_init, obj_rot_init=rot_init,
                                  on_mini_batch=save_neural, **self.loop_kw)
            smpl = out["smpl"]
            p = torch.cat([smpl.global_pose, smpl.body_pose, smpl.hand_pose], 1)
            b = torch.cat([smpl.top_betas, smpl.other_betas], 1)
            vio.save_smpl_params(folders, args.test_kid, p, b, smpl.trans)                      # save_outputs (recon_fit_base.py:292-313)
            vio.save_object_params(folders, args.test_kid, out["obj_R"], out["obj_t"], out["obj_s"])
            dt = time.time() - start_time
            print(f"Optimization time for one batch of size {len(batch)}: {dt:.5f} seconds, avg={dt / len(batch):.5f}")
            done += len(batch)
        return done


def recon_fit(args):
    assert args.triplane_type != 'gt', 'do not use gt as triplane!'
    fitter = ReconFitterTriVisFull(args.seq_folder, debug=args.display, outpath=args.outpath, args=args)
    fitter.fit_recon(args)
    print('all done')


def cli(argv=None) -> int:
    """recon_fit_trivis_full.py:488-500; the reference prints the traceback and exits 0 -- here the failure is also reported to the shell."""
    import traceback
    parser = ReconFitterTriVisFull.get_parser()
    args = parser.parse_args(argv)
    configs = ReconFitterTriVisFull.merge_configs(args, load_configs(args.exp_name))
    try:
        recon_fit(configs)
    except Exception:
        print(traceback.format_exc())
        return 1
    return 0


if __name__ == '__main__':
    sys.exit(cli())
