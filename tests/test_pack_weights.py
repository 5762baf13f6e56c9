"""vt_pack_weights_* (csrc/pack.cu, host C behind the C ABI) against an independent torch statement of every kernel layout -- CPU only,
no compute on a GPU.  The torch functions below are the round-1 packers (then the product path, now the checker): tensor permutes /
transposes / index_select, sharing no code with the C loops."""
from typing import Dict, Tuple

import numpy as np
import pytest
import torch

from vistracker_b200 import weights as W

LO_SCALE = 2048.0
MMA_KC = 64


def split_f16(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """fp32 -> (hi, lo) fp16 planes with x ~= hi + lo * 2^-11 (same rounding as the device-side split)."""
    x = x.float()
    if float(x.abs().max()) > 65504.0:
        raise ValueError("weight magnitude exceeds the fp16 range; the fp16x2 tensor-core path cannot represent it")
    hi = x.half()
    lo = ((x - hi.float()) * LO_SCALE).half()
    return hi.contiguous(), lo.contiguous()


def ref_pack_conv(w: torch.Tensor) -> Dict[str, torch.Tensor]:
    """Conv2d weight [Cout, Cin, k, k] -> {'ffma': fp32 [k*k, Cin, Cout], 'hi'/'lo': fp16 [k*k, Cout, Cin_pad]}."""
    cout, cin, kh, kw = w.shape
    assert kh == kw
    taps = w.permute(2, 3, 0, 1).reshape(kh * kw, cout, cin).float()          # [tap, Cout, Cin]
    cin_pad = (cin + MMA_KC - 1) // MMA_KC * MMA_KC
    padded = torch.zeros(kh * kw, cout, cin_pad, dtype=torch.float32, device=w.device)
    padded[:, :, :cin] = taps
    hi, lo = split_f16(padded)
    return {"ffma": taps.transpose(1, 2).contiguous(), "hi": hi, "lo": lo, "ks": kh, "cin": cin, "cout": cout,
            "cin_pad": cin_pad}


def ref_pack_stem(w: torch.Tensor) -> torch.Tensor:
    """Conv2d(cin, cout, 7, stride 2) weight [Cout, Cin, 7, 7] -> fp32 [49*Cin, Cout] (tap-major, then input channel)."""
    cout, cin = w.shape[:2]
    return w.permute(2, 3, 1, 0).reshape(49 * cin, cout).float().contiguous()


# ---- decoder packing (csrc/query.cu) ---------------------------------------------------------------------------
Q_K, Q_H = 616, 128
HEADS = ("df", "pca_predictor", "part_predictor", "center_predictor", "visib_predictor")   # output order df|pca|parts|centers|vis
HEAD_NOUT = (2, 9, 14, 3, 1)


def feature_permutation(c_im=256, c_tmpx=64, c_tt=32, c_tf=64) -> torch.Tensor:
    """internal index -> reference feature index.  Reference order (model/chore_triplane.py:139-151):
    im_feat | x,y,z-2.2 | tmpx | tri_tmpx r,b,t | tri_feat r | b | t ; internal order moves the 3 scalars to the end."""
    n_rest = c_tmpx + 3 * c_tt + 3 * c_tf
    ref = list(range(c_im)) + [c_im + 3 + i for i in range(n_rest)] + [c_im, c_im + 1, c_im + 2]
    return torch.tensor(ref, dtype=torch.long)


def ref_pack_decoders(sd: Dict[str, torch.Tensor], device) -> torch.Tensor:
    """Five Conv1d(k=1) MLPs (model/chore.py:113-126) -> one fp32 buffer, k-major, first layer rows permuted/padded."""
    perm = feature_permutation()
    chunks = []
    for name, nout in zip(HEADS, HEAD_NOUT):
        w1 = sd[f"{name}.0.weight"][:, :, 0].float()                 # [128, 611]
        assert w1.shape == (Q_H, perm.numel()), f"{name}.0.weight has shape {tuple(w1.shape)}"
        w1p = torch.zeros(Q_K, Q_H)
        w1p[: perm.numel()] = w1[:, perm].t().cpu()
        chunks += [w1p.reshape(-1), sd[f"{name}.0.bias"].float().cpu()]
        for idx in (2, 4):
            chunks += [sd[f"{name}.{idx}.weight"][:, :, 0].float().t().contiguous().cpu().reshape(-1),
                       sd[f"{name}.{idx}.bias"].float().cpu()]
        w4 = torch.zeros(Q_H, 16)
        w4[:, :nout] = sd[f"{name}.6.weight"][:, :, 0].float().t().cpu()
        b4 = torch.zeros(16)
        b4[:nout] = sd[f"{name}.6.bias"].float().cpu()
        chunks += [w4.reshape(-1), b4]
    return torch.cat(chunks).contiguous().to(device)


Q_KB = 640


def ref_pack_decoders_bwd(sd: Dict[str, torch.Tensor], device) -> torch.Tensor:
    """Transposed copies for the analytic backward pass (csrc/query.cu): per head W1b [128][640] (feature columns in the
    internal order, zero padded), W2b, W3b [128][128] and W4b [16][128] -- i.e. the torch [out][in] layout."""
    perm = feature_permutation()
    chunks = []
    for name, nout in zip(HEADS, HEAD_NOUT):
        w1b = torch.zeros(Q_H, Q_KB)
        w1b[:, : perm.numel()] = sd[f"{name}.0.weight"][:, :, 0].float().cpu()[:, perm]
        chunks.append(w1b.reshape(-1))
        for idx in (2, 4):
            chunks.append(sd[f"{name}.{idx}.weight"][:, :, 0].float().cpu().contiguous().reshape(-1))
        w4b = torch.zeros(16, Q_H)
        w4b[:nout] = sd[f"{name}.6.weight"][:, :, 0].float().cpu()
        chunks.append(w4b.reshape(-1))
    return torch.cat(chunks).contiguous().to(device)


def feature_permutation_tc() -> torch.Tensor:
    """tensor-core kernel feature index (csrc/query_tc.cu, 640 padded) -> reference feature index, -1 for zero padding:
    im_feat 256 | tmpx 64 | tri_feat r, b, t 3x64 | tri_tmpx r 32, b 32 | tri_tmpx t 32, x, y, z-2.2, 29 zeros."""
    ref = list(range(256)) + list(range(259, 323)) + list(range(419, 611)) + list(range(323, 387)) + list(range(387, 419)) + [256, 257, 258]
    return torch.tensor(ref + [-1] * (640 - len(ref)), dtype=torch.long)


def split_f16_unscaled(x: torch.Tensor):
    """x ~= hi + lo with lo NOT rescaled (single-accumulator scheme of csrc/query_tc.cu)."""
    x = x.float()
    if float(x.abs().max()) > 65504.0:
        raise ValueError("weight magnitude exceeds the fp16 range")
    hi = x.half()
    return hi.contiguous(), (x - hi.float()).half().contiguous()


def ref_pack_decoders_tc(sd: Dict[str, torch.Tensor], device):
    """fp16 hi/lo planes for the tcgen05 decoder kernel: W1 [5*128, 640] (rows = head, unit; columns in the kernel's feature
    order) and W2|W3 [2*5*128, 128] (layer-major, then head) -- torch's [out][in] layout, i.e. K-major B operands."""
    perm = feature_permutation_tc()
    w1 = torch.zeros(5 * Q_H, 640)
    w23 = torch.zeros(2 * 5 * Q_H, Q_H)
    valid = perm >= 0
    for h, name in enumerate(HEADS):
        w = sd[f"{name}.0.weight"][:, :, 0].float().cpu()
        w1[h * Q_H:(h + 1) * Q_H][:, valid] = w[:, perm[valid]]
        for li, idx in enumerate((2, 4)):
            w23[(li * 5 + h) * Q_H:(li * 5 + h + 1) * Q_H] = sd[f"{name}.{idx}.weight"][:, :, 0].float().cpu()
    w1h, w1l = split_f16_unscaled(w1)
    w2h, w2l = split_f16_unscaled(w23)
    return tuple(t.to(device) for t in (w1h, w1l, w2h, w2l))


def ref_pack_decoders_tc_bwd(sd: Dict[str, torch.Tensor], device):
    """Transposed fp16 hi/lo planes for the tcgen05 backward kernel (csrc/query_bwd_tc.cu): W2^T | W3^T [2*5*128, 128] (rows = input
    unit j, columns = output unit k: the B operand of g_in[j] = sum_k g_out[k] W[k][j]) and W1^T [5*640, 128] (rows = feature in the
    kernel's order, zero rows for the padding)."""
    perm = feature_permutation_tc()
    valid = perm >= 0
    w23t = torch.zeros(2 * 5 * Q_H, Q_H)
    w1t = torch.zeros(5 * 640, Q_H)
    for h, name in enumerate(HEADS):
        w = sd[f"{name}.0.weight"][:, :, 0].float().cpu()                     # [128 out, 611 in]
        blk = torch.zeros(640, Q_H)
        blk[valid] = w[:, perm[valid]].t()
        w1t[h * 640:(h + 1) * 640] = blk
        for li, idx in enumerate((2, 4)):
            w23t[(li * 5 + h) * Q_H:(li * 5 + h + 1) * Q_H] = sd[f"{name}.{idx}.weight"][:, :, 0].float().cpu().t()
    a, b = split_f16_unscaled(w23t)
    c, d = split_f16_unscaled(w1t)
    return tuple(t.to(device) for t in (a, b, c, d))


# ================================================================================================================== tests
def _decoder_sd(seed=0):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, nout in zip(HEADS, HEAD_NOUT):
        dims = [(128, 611), (128, 128), (128, 128), (nout, 128)]
        for idx, (o, i) in zip((0, 2, 4, 6), dims):
            sd[f"{name}.{idx}.weight"] = torch.randn(o, i, 1, generator=g) * 0.2
            sd[f"{name}.{idx}.bias"] = torch.randn(o, generator=g) * 0.1
    return sd


@pytest.mark.parametrize("cout,cin,ks", [(64, 64, 3), (128, 37, 1), (32, 8, 3), (256, 256, 1), (5, 3, 7)])
def test_conv_pack_equals_the_torch_statement(cout, cin, ks):
    w = torch.randn(cout, cin, ks, ks, generator=torch.Generator().manual_seed(cout + cin)) * 0.3
    w[0, 0, 0, 0] = 0.0; w[-1, -1, -1, -1] = 1e-7                   # lo plane of tiny / zero values
    got, ref = W.pack_conv(w), ref_pack_conv(w)
    assert got["cin_pad"] == ref["cin_pad"] and got["ks"] == ks
    for k in ("ffma", "hi", "lo"):
        assert got[k].dtype == ref[k].dtype and got[k].shape == ref[k].shape and torch.equal(got[k], ref[k]), k


def test_conv_pack_rejects_values_beyond_fp16():
    w = torch.zeros(8, 8, 1, 1); w[3, 2] = 7.0e4
    with pytest.raises(ValueError):
        W.pack_conv(w)


def test_stem_pack_equals_the_torch_statement():
    w = torch.randn(64, 3, 7, 7, generator=torch.Generator().manual_seed(2))
    assert torch.equal(W.pack_stem(w), ref_pack_stem(w))


def test_decoder_packs_equal_the_torch_statement():
    sd = _decoder_sd()
    assert torch.equal(W.pack_decoders(sd, "cpu"), ref_pack_decoders(sd, "cpu"))
    assert torch.equal(W.pack_decoders_bwd(sd, "cpu"), ref_pack_decoders_bwd(sd, "cpu"))
    for a, b in zip(W.pack_decoders_tc(sd, "cpu"), ref_pack_decoders_tc(sd, "cpu")):
        assert a.dtype == torch.float16 and a.shape == b.shape and torch.equal(a, b)
    for a, b in zip(W.pack_decoders_tc_bwd(sd, "cpu"), ref_pack_decoders_tc_bwd(sd, "cpu")):
        assert a.dtype == torch.float16 and a.shape == b.shape and torch.equal(a, b)


def test_decoder_pack_checks_the_checkpoint_shapes():
    sd = _decoder_sd()
    sd["df.0.weight"] = torch.zeros(128, 600, 1)
    with pytest.raises(AssertionError):
        W.pack_decoders(sd, "cpu")


def test_smpl_pack_equals_the_torch_statement():
    """vt_pack_weights_smpl against the einsum / argsort statement SMPL_Layer.from_buffers used in round 1."""
    import ctypes
    from vistracker_b200 import _lib
    rng = np.random.default_rng(3)
    V, J, nb = 211, 7, 10
    vt, sd, pd = rng.standard_normal((V, 3)), rng.standard_normal((V, 3, nb)) * 0.1, rng.standard_normal((V, 3, 9 * (J - 1))) * 0.01
    jr = np.abs(rng.standard_normal((J, V))); jr[jr < 1.2] = 0; jr /= jr.sum(1, keepdims=True)
    w = np.zeros((V, J), np.float32)
    for v in range(V):
        idx = rng.choice(J, size=int(rng.integers(1, 5)), replace=False)
        w[v, idx] = rng.random(len(idx)).astype(np.float32) + 0.05
    w /= w.sum(1, keepdims=True)
    parents = np.array([-1, 0, 0, 1, 2, 3, 3], np.int32)
    dims = [ctypes.c_int() for _ in range(4)]
    _lib.call("vt_smpl_pack_dims", V, J, nb, w.ctypes.data_as(ctypes.c_void_p), *(ctypes.byref(d) for d in dims))
    kd, kdp, nv3p, nnz = (d.value for d in dims)
    assert kd == 9 * (J - 1) + nb and kdp == (kd + 3) // 4 * 4 and nv3p == (3 * V + 3) // 4 * 4 and nnz == int((w != 0).sum(1).max())
    out = dict(templ=np.empty(3 * V, np.float32), dirs=np.empty((kdp, nv3p), np.float32), dirsT=np.empty((nv3p, kdp), np.float32),
               j_templ=np.empty((J, 3), np.float32), j_dirs=np.empty((J, 3, nb), np.float32), parents=np.empty(J, np.int32),
               skin_idx=np.empty((V, nnz), np.int32), skin_w=np.empty((V, nnz), np.float32))
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    _lib.call("vt_pack_weights_smpl", P(vt), P(sd), P(pd), P(jr), P(w), P(parents), V, J, nb, *(P(out[k]) for k in
                                                                                          ("templ", "dirs", "dirsT", "j_templ", "j_dirs", "parents", "skin_idx", "skin_w")))
    tvt, tsd, tpd, tjr, tw = (torch.from_numpy(a) for a in (vt, sd, pd, jr, w))
    dirs = torch.zeros(kdp, nv3p)
    dirs[:9 * (J - 1), :3 * V] = tpd.reshape(3 * V, -1).t().float()
    dirs[9 * (J - 1):kd, :3 * V] = tsd.reshape(3 * V, -1).t().float()
    order = torch.argsort((tw != 0).to(torch.int8), dim=1, descending=True, stable=True)[:, :nnz]
    assert np.array_equal(out["templ"], tvt.reshape(-1).float().numpy())
    assert np.array_equal(out["dirs"], dirs.numpy()) and np.array_equal(out["dirsT"], dirs.t().contiguous().numpy())
    assert np.allclose(out["j_templ"], (tjr @ tvt).float().numpy(), rtol=0, atol=1e-7)
    assert np.allclose(out["j_dirs"], torch.einsum("jv,vck->jck", tjr, tsd).float().numpy(), rtol=0, atol=1e-7)
    assert out["parents"].tolist() == [0, 0, 0, 1, 2, 3, 3]
    assert np.array_equal(out["skin_idx"], order.numpy().astype(np.int32)) and np.array_equal(out["skin_w"], torch.gather(tw, 1, order).numpy())


def test_workspace_sizes():
    from vistracker_b200 import _lib
    lib = _lib.load()
    assert lib.vt_workspace_bytes_procrustes(5) == 5 * 16 * 8 and lib.vt_workspace_bytes_procrustes(0) == 0
    assert lib.vt_workspace_bytes_raster_cull(3, 800) == 4 * lib.vt_raster_cull_floats(3, 800)
