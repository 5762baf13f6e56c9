"""ctypes binding of libvistracker_sm100a.so (the C ABI declared in include/vistracker_b200.h).

There is no fallback: if the library is missing or a call is rejected, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvistracker_sm100a.so")

_p, _i, _f, _ll, _d = C.c_void_p, C.c_int, C.c_float, C.c_longlong, C.c_double

# name -> (restype, argtypes); must list every symbol of include/vistracker_b200.h (tests/test_cabi.py checks both ways)
SIGNATURES = {
    "vt_last_error": (C.c_char_p, []),
    "vt_version": (_i, []),
    "vt_compiled_arch": (_i, []),
    "vt_conv_cin_pad": (_i, [_i]),
    "vt_pack_weights_conv": (_i, [_p, _i, _i, _i, _p, _p, _p]),
    "vt_pack_weights_stem": (_i, [_p, _i, _i, _p]),
    "vt_pack_weights_decoders": (_i, [_p, _p, _p, _p]),
    "vt_pack_weights_decoders_tc": (_i, [_p] * 9),
    "vt_smpl_pack_dims": (_i, [_i, _i, _i, _p, _p, _p, _p, _p]),
    "vt_pack_weights_smpl": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p]),
    "vt_workspace_bytes_raster_cull": (_ll, [_i, _i]),
    "vt_workspace_bytes_procrustes": (_ll, [_i]),
    "vt_stem_conv7x7s2": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _p, _p, _i, _p, _p, _i, _p]),
    "vt_gn_finalize": (_i, [_p, _i, _p, _p, _i, _i, _i, _ll, _f, _p, _p, _p]),
    "vt_affine_act": (_i, [_p, _i, _p, _p, _i, _i, _i, _i, _p, _i, _p, _i, _p]),
    "vt_prep_split": (_i, [_p, _i, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "vt_prep_split_gn": (_i, [_p, _i, _p, _i, _p, _p, _i, _ll, _f, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "vt_conv_mma": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _p, _p, _i, _p, _p, _i, _p, _i, _p, _i, _p]),
    "vt_conv_mma_dual": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _p, _p, _i, _p, _p, _i, _p, _i, _p, _i, _p, _i, _p, _i, _p, _i, _p]),
    "vt_conv_ffma": (_i, [_p, _i, _p, _p, _i, _i, _i, _i, _i, _i, _p, _i, _p, _p, _i, _p, _i, _p, _i, _p]),
    "vt_add": (_i, [_p, _i, _p, _i, _i, _i, _i, _p, _i, _p, _i, _p]),
    "vt_avgpool2": (_i, [_p, _i, _i, _i, _i, _p, _p, _i, _p]),
    "vt_upsample2x_add": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _i, _p]),
    "vt_query_wpack_floats": (_ll, []),
    "vt_query_fwd": (_i, [_p, _p, _p, _i, _i, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p]),
    "vt_query_fwd_tc": (_i, [_p, _p, _p, _i, _i, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "vt_query_bwd_tc": (_i, [_p, _p, _p, _i, _i, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p] + [_p] * 8 + [_p, _i, _p, _p, _p]),
    "vt_query_project_step_tc": (_i, [_p, _p, _p, _i, _i, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p] + [_p] * 8 + [_i, _f, _p, _p, _p, _p]),
    "vt_query_fwd_tc_heads": (_i, [_p, _p, _p, _i, _i, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _i, _p, _p, _p, _p]),
    "vt_query_losses_tc": (_i, [_p, _p, _p, _i, _i, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p] + [_p] * 8 + [_i, _f, _p, _p, _p, _p, _p, _i, _p, _p, _p]),
    "vt_query_losses_merged_tc": (_i, [_p, _p, _p, _i, _i, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p] + [_p] * 8 + [_i, _f, _p, _p, _f, _p, _f, _p, _p, _p, _p, _p]),
    "vt_query_wpack_bwd_floats": (_ll, []),
    "vt_query_bwd": (_i, [_p, _p, _p, _i, _i, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p]),
}



class SmplModelStruct(C.Structure):
    """struct vt_smpl_model (include/vistracker_b200.h)."""
    _fields_ = [("V", _i), ("J", _i), ("n_betas", _i), ("kd", _i), ("kdp", _i), ("nv3p", _i), ("nnz", _i),
                ("templ", _p), ("dirs", _p), ("dirsT", _p), ("j_templ", _p), ("j_dirs", _p), ("parents", _p),
                ("skin_idx", _p), ("skin_w", _p)]


_ms = C.POINTER(SmplModelStruct)
SIGNATURES.update({
    "vt_smpl_fwd": (_i, [_ms, _p, _p, _p, _p, _f, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "vt_smpl_bwd": (_i, [_ms, _p, _p, _p, _p, _p, _p, _p, _p, _f, _i, _p, _p, _p, _p, _p, _p, _p, _p]),
    "vt_landmarks_fwd": (_i, [_p, _i, _i, _p, _p, _p, _i, _p, _p]),
    "vt_landmarks_bwd": (_i, [_p, _i, _i, _p, _p, _p, _i, _p, _p]),
    "vt_fit_ctrl_words": (_i, []),
    "vt_fit_begin_step": (_i, [_p, _p]),
    "vt_fit_kpts": (_i, [_p, _p, _i, _i, _p, _p, _p, _p, _p]),
    "vt_fit_temporal_verts": (_i, [_p, _i, _i, _p, _p, _p, _p]),
    "vt_fit_pose_terms": (_i, [_p, _p, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "vt_fit_adam": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _p, _p]),
    "vt_fit_end_step": (_i, [_p, _i, _i, _p, _p, _i, _p]),
    "vt_so3_project_fwd": (_i, [_p, _i, _p, _p]),
    "vt_so3_project_bwd": (_i, [_p, _p, _i, _p, _p]),
    "vt_pca_orientation": (_i, [_p, _p, _i, _p, _i, _p, _p]),
    "vt_chamfer_fwd": (_i, [_p, _p, _p, _p, _i, _p, _p, _p, _p]),
    "vt_chamfer_bwd": (_i, [_p, _p, _p, _p, _i, _p, _p, _p, _p, _p, _p]),
    "vt_query_bwd_heads": (_i, [_p, _p, _p, _i, _i, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _i, _p, _p]),
    "vt_query_project_step": (_i, [_p, _p, _p, _i, _i, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _i, _f, _p, _p, _p, _p]),
    "vt_procrustes": (_i, [_p, _p, _i, _i, _p, _p, _p, _p, _p]),
    "vt_similarity_apply": (_i, [_p, _i, _i, _p, _p, _p, _p, _p]),
    "vt_nn_dist": (_i, [_p, _i, _p, _i, _i, _p, _p, _p]),
    "vt_smoothnet_pack_floats": (_ll, [_i]),
    "vt_smooth_pack_smplt": (_i, [_p, _i, _p, _p, _i, _p, _p]),
    "vt_smoothnet_clips": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p]),
    "vt_smooth_window_mean": (_i, [_p, _p, _i, _i, _i, _i, _i, _p, _p]),
    "vt_smooth_unpack_smplt": (_i, [_p, _i, _p, _p, _p, _p]),
    "vt_smooth_rot6d_to_rotmat": (_i, [_p, _i, _i, _p, _p]),
    "vt_infill_layer_pack_floats": (_ll, [_i, _i]),
    "vt_infill_head": (_i, [_p, _i, _i, _p, _p, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "vt_infill_attn": (_i, [_p, _p, _i, _i, _i, _i, _p, _p]),
    "vt_infill_tail": (_i, [_p, _i, _p, _i, _i, _i, _i, _i, _i, _p, _p, _p, _i, _p, _i, _p, _p, _p]),
    "vt_infill_mlp": (_i, [_p, _i, _i, _i, _p, _p, _p, _i, _p]),
    "vt_infill_pack_clip": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p]),
    "vt_infill_commit_clip": (_i, [_p, _i, _i, _i, _i, _p, _p]),
    "vt_mask_bbox": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _p]),
    "vt_prepare_image_crop": (_i, [_p, _p, _p, _p, _i, _i, _i, _p, _i, _i, _p, _p, _i, _p]),
    "vt_raster_cull_floats": (_ll, [_i, _i]),
    "vt_raster_fwd": (_i, [_p, _p, _i, _i, _i, _i, _p, _i, _p, _p, _p, _p, _p, _p]),
    "vt_raster_bwd": (_i, [_p, _p, _i, _i, _i, _i, _p, _i, _p, _p, _p, _p, _p, _p, _p]),
    "vt_workspace_bytes_raster_bwd": (_ll, [_i, _i]),
    "vt_raster_bwd_ws": (_i, [_p, _p, _i, _i, _i, _i, _p, _i, _p, _p, _p, _p, _p, _p, _p, _p]),
    "vt_recon_ctrl_words": (_i, []),
    "vt_recon_hist_ld": (_i, []),
    "vt_zero": (_i, [_p, _ll, _p]),
    "vt_recon_point_terms": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _i, _i, _f, _p, _p, _i, _i, _f, _p, _p, _p, _p]),
    "vt_recon_kpts": (_i, [_p, _p, _p, _i, _i, _p, _p, _p, _p, _p]),
    "vt_recon_pose_terms": (_i, [_p, _p, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "vt_recon_adam_smpl": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _p, _p]),
    "vt_recon_adam_obj": (_i, [_p, _p, _p, _p, _p, _p, _i, _p, _p]),
    "vt_recon_end_step": (_i, [_p, _i, _p, _i, _p, _p, _p, _i, _p]),
    "vt_recon_obj_noise": (_i, [_p, _p, _i, _p, _p, _p, _p]),
    "vt_recon_obj_transform": (_i, [_p, _i, _p, _p, _p, _i, _i, _p, _p]),
    "vt_recon_obj_transform_bwd": (_i, [_p, _i, _p, _p, _i, _i, _i, _p, _p, _p]),
    "vt_recon_sil_loss": (_i, [_p, _p, _p, _p, _i, _i, _p, _p, _p, _p]),
    "vt_recon_obj_small_terms": (_i, [_p, _p, _p, _f, _i, _i, _p, _p, _p, _p]),
    "vt_recon_gather_rows": (_i, [_p, _p, _i, _p, _p]),
    "vt_recon_scatter_add_rows": (_i, [_p, _p, _i, _p, _p]),
})

_lib = None


def load():
    """Load the shared library (building is the job of __graft_entry__.build / vistracker_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -m vistracker_b200.build` (there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def call(name: str, *args):
    """Invoke an int-returning entry point and raise on rejection."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {lib.vt_last_error().decode()}")


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


# ---- NVTX ranges around the stages of a batch (VT_NVTX=1; nsys / ncu --nvtx pick them up).  A no-op context manager otherwise.
import contextlib as _contextlib

_NVTX = os.environ.get("VT_NVTX", "0") not in ("", "0")


@_contextlib.contextmanager
def nvtx_range(name: str):
    if not _NVTX:
        yield
        return
    import torch
    torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        torch.cuda.nvtx.range_pop()
