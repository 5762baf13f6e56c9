"""SmoothNet stage on B200: drop-ins for the numeric core of ``smoothnet/smooth_smplt.py`` and
``smoothnet/smooth_objrot.py``.

The reference loads one pickle per frame, runs ``SmoothNetSMPL`` / ``SmoothNet`` (smoothnet/models) on CPU-built sliding windows and
writes a packed joblib file that the next script reads back.  Here the trajectory is the device tensor the fitting stage produced
(``parallel.gather_trajectory``: [T, 169] = pose 156 | betas 10 | trans 3) and every step is a kernel of libvistracker_sm100a.so:
rotation conversion, the fused window-gather + MLP (``vt_smoothnet_clips``), the window mean, the conversion back.  File IO
(``load_inputs_raw`` / ``dump_packed``) stays with the caller.
"""
from __future__ import annotations

from typing import Dict

import torch

from . import _lib

WINDOW, HIDDEN, RES_HIDDEN = 64, 512, 16          # smoothnet/configs/pw3d_spin_3D.yaml, --slide_window_size default 64
P, S = _lib.ptr, _lib.stream_ptr


def pack_smoothnet(sd: Dict[str, torch.Tensor], prefix: str, device) -> tuple:
    """k-major fp32 pack of one ``SmoothNet`` (keys ``<prefix>encoder.0.*``, ``<prefix>res_blocks.<i>.linear{1,2}.*``,
    ``<prefix>decoder.*``): (buffer, number of residual blocks)."""
    def w(name):
        return sd[prefix + name].float().cpu()
    We = w("encoder.0.weight")
    if tuple(We.shape) != (HIDDEN, WINDOW):
        raise RuntimeError(f"{prefix}encoder.0.weight has shape {tuple(We.shape)}; the kernel is built for ({HIDDEN}, {WINDOW})")
    chunks = [We.t().contiguous().reshape(-1), w("encoder.0.bias")]
    n_blocks = 0
    while f"{prefix}res_blocks.{n_blocks}.linear1.weight" in sd:
        W1, W2 = w(f"res_blocks.{n_blocks}.linear1.weight"), w(f"res_blocks.{n_blocks}.linear2.weight")
        if tuple(W1.shape) != (RES_HIDDEN, HIDDEN) or tuple(W2.shape) != (HIDDEN, RES_HIDDEN):
            raise RuntimeError(f"{prefix}res_blocks.{n_blocks}: unsupported shapes {tuple(W1.shape)} / {tuple(W2.shape)}")
        chunks += [W1.t().contiguous().reshape(-1), w(f"res_blocks.{n_blocks}.linear1.bias"),
                   W2.t().contiguous().reshape(-1), w(f"res_blocks.{n_blocks}.linear2.bias")]
        n_blocks += 1
    Wd = w("decoder.weight")
    if tuple(Wd.shape) != (WINDOW, HIDDEN):
        raise RuntimeError(f"{prefix}decoder.weight has shape {tuple(Wd.shape)}")
    chunks += [Wd.t().contiguous().reshape(-1), w("decoder.bias")]
    buf = torch.cat(chunks).contiguous().to(device)
    assert buf.numel() == _lib.load().vt_smoothnet_pack_floats(n_blocks)
    return buf, n_blocks


class _SmootherBase:
    def __init__(self, device=None):
        self.device = torch.device(device) if device is not None else torch.device("cuda", 0)
        if self.device.type != "cuda":
            raise RuntimeError("vistracker_b200 has no CPU path: the smoothers need a CUDA device")
        _lib.load()

    def _clips(self, seq, c0, nC, relative, pack, clips):
        wpack, n_blocks = pack
        L, D = seq.shape
        _lib.call("vt_smoothnet_clips", P(seq), L, D, c0, nC, int(relative), WINDOW, HIDDEN, RES_HIDDEN, n_blocks, P(wpack), P(clips), S())

    def _mean(self, clips, seq, pass0, passN):
        L, D = seq.shape
        out = torch.empty_like(seq)
        _lib.call("vt_smooth_window_mean", P(clips), P(seq), L, D, WINDOW, pass0, passN, P(out), S())
        return out


class SMPLTSmoother(_SmootherBase):
    """``SMPLTSmoother`` (smoothnet/smooth_smplt.py) with a ``SmoothNetSMPL`` checkpoint (keys ``pose_net.*`` / ``trans_net.*``)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device=None):
        super().__init__(device)
        sd = {(k[7:] if k.startswith("module.") else k): v for k, v in state_dict.items()}
        with torch.cuda.device(self.device):
            self.pose_net = pack_smoothnet(sd, "pose_net.", self.device)
            self.trans_net = pack_smoothnet(sd, "trans_net.", self.device)

    def smooth(self, poses: torch.Tensor, betas: torch.Tensor, trans: torch.Tensor) -> Dict[str, torch.Tensor]:
        """poses [T, 72 | 156] axis-angle, betas [T, 10], trans [T, 3] -> {'poses' [T, 72], 'betas' [T, 10], 'trans' [T, 3]} (device
        tensors; preprocess_input -> model -> post_processing of the reference, T >= 64)."""
        T = poses.shape[0]
        if T < WINDOW:
            raise ValueError(f"SmoothNet needs at least one window of {WINDOW} frames, got {T}")
        dev = self.device
        poses, betas, trans = (t.to(dev, torch.float32).contiguous() for t in (poses, betas, trans))
        with torch.cuda.device(dev):
            seq = torch.empty(T, 157, device=dev)
            _lib.call("vt_smooth_pack_smplt", P(poses), poses.shape[1], P(betas), P(trans), T, P(seq), S())
            clips = torch.empty(T - WINDOW + 1, WINDOW, 157, device=dev)
            self._clips(seq, 0, 144, False, self.pose_net, clips)
            self._clips(seq, 154, 3, True, self.trans_net, clips)
            den = self._mean(clips, seq, 144, 10)
            out_p, out_b, out_t = torch.empty(T, 72, device=dev), torch.empty(T, 10, device=dev), torch.empty(T, 3, device=dev)
            _lib.call("vt_smooth_unpack_smplt", P(den), T, P(out_p), P(out_b), P(out_t), S())
        return {"poses": out_p, "betas": out_b, "trans": out_t, "rot6d": den}

    def smooth_trajectory(self, traj: torch.Tensor) -> Dict[str, torch.Tensor]:
        """traj [T, 169] = pose 156 | betas 10 | trans 3, the block ``parallel.gather_trajectory`` returns."""
        return self.smooth(traj[:, :156], traj[:, 156:166], traj[:, 166:169])


class ObjrotSmoother(_SmootherBase):
    """``ObjrotSmoother`` (smoothnet/smooth_objrot.py) with a plain ``SmoothNet`` checkpoint."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device=None):
        super().__init__(device)
        sd = {(k[7:] if k.startswith("module.") else k): v for k, v in state_dict.items()}
        with torch.cuda.device(self.device):
            self.net = pack_smoothnet(sd, "", self.device)

    def smooth(self, rot: torch.Tensor) -> torch.Tensor:
        """rot [T, 3, 3] rotation matrices ('obj_rot' of load_inputs_raw, i.e. already the real rotations) -> 'obj_angles' [T, 3, 3]
        (the smoothed rotations, transposed as the reference stores them)."""
        T = rot.shape[0]
        if T < WINDOW:
            raise ValueError(f"SmoothNet needs at least one window of {WINDOW} frames, got {T}")
        dev = self.device
        seq = rot.to(dev, torch.float32).reshape(T, 3, 3)[:, :, :2].reshape(T, 6).contiguous()       # rotmat_to_6d
        with torch.cuda.device(dev):
            clips = torch.empty(T - WINDOW + 1, WINDOW, 6, device=dev)
            self._clips(seq, 0, 6, False, self.net, clips)
            den = self._mean(clips, seq, 0, 0)
            out = torch.empty(T, 3, 3, device=dev)
            _lib.call("vt_smooth_rot6d_to_rotmat", P(den), T, 1, P(out), S())
        return out
