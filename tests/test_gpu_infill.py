"""HVOP-Net kernels (vt_infill_*) through vistracker_b200.infill against the reference's outputs (tests/golden/infill_small.npz) and the
float64 restatement (oracle/infill_ref.py)."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import infill_ref as R
from vistracker_b200.synth import synthetic_infill_sequence, synthetic_infill_state_dict

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "infill_small.npz")


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def _model(opt, sd):
    from vistracker_b200.infill import ConditionalMInfiller
    return ConditionalMInfiller(opt, device="cuda:0").eval().load_state_dict(sd)


def _gold():
    g = np.load(GOLD)
    opt = json.loads(str(g["opt_json"]))
    return g, opt, synthetic_infill_state_dict(opt, seed=21)


def test_forward_matches_reference_golden():
    _need_gpu()
    g, opt, sd = _gold()
    net = _model(opt, sd)
    for tag in ("a", "b"):
        pred = net(g[f"{tag}_data_smpl"], g[f"{tag}_mask_smpl"], g[f"{tag}_data_obj"], g[f"{tag}_mask_obj"])
        assert pred.shape == g[f"{tag}_pred"].shape
        assert rel_err(pred.cpu(), g[f"{tag}_pred"]) < 2e-5, tag                   # fp32 both sides; the float64 oracle sits 1e-6 from either


def test_forward_other_widths_against_oracle():
    """Odd widths, a closing LayerNorm (pre_norm option), relu / leaky_relu, a deeper predictor, ragged token blocks (T = 37, B = 3)."""
    _need_gpu()
    opt = dict(clip_len=37, dim_smpl=21, dim_obj=6, out_dim=6, num_layers_smpl=1, d_model_smpl=48, num_heads_smpl=3, dim_forward_smpl=72,
               pre_norm_smpl=True, activation_smpl="relu", num_layers_obj=3, d_model_obj=20, num_heads_obj=5, dim_forward_obj=40, pre_norm_obj=False,
               activation_obj="leaky_relu", num_layers_joint=2, num_heads_joint=2, dim_forward_joint=100, pre_norm_joint=True, activation_joint="gelu",
               hidden_dims=[24, 12])
    sd = synthetic_infill_state_dict(opt, seed=4)
    net = _model(opt, sd)
    gen = torch.Generator().manual_seed(9)
    ds, do = torch.randn(3, 37, 21, generator=gen), torch.randn(3, 37, 6, generator=gen)
    mo = torch.rand(3, 37, generator=gen) < 0.5
    mo[:, 0] = False
    ms = torch.zeros(3, 37, dtype=torch.bool)
    ref = R.cond_infiller_forward(sd, opt, ds, ms, do, mo)
    got = net(ds, ms, do, mo)
    assert rel_err(got.cpu(), ref) < 2e-5


def test_fully_masked_clip_is_nan_like_torch():
    _need_gpu()
    g, opt, sd = _gold()
    net = _model(opt, sd)
    mo = np.ones((1, 47), bool)
    got = net(g["b_data_smpl"], g["b_mask_smpl"] * False, g["b_data_obj"], mo)
    assert bool(torch.isnan(got).all())                                            # softmax over an empty key set (nn.MultiheadAttention does the same)


def test_autoregressive_sequence_matches_reference_loop():
    _need_gpu()
    from vistracker_b200.infill import CondMotionInfillAutoreg
    g, opt, sd = _gold()
    net = _model(opt, sd)
    L = int(g["seq_L"])
    seq = synthetic_infill_sequence(L, seed=5)
    outs = []
    eager, drv = CondMotionInfillAutoreg(net, use_graph=False), CondMotionInfillAutoreg(net, use_graph=True)
    for d in (eager, drv, drv):                                                    # eager launches, graph capture + replay, replay
        res = d.infill(*seq, occ_thres=0.5)
        outs.append(res["obj_angles"].cpu().numpy())
        assert np.abs(outs[-1] - g["seq_obj_angles"]).max() < 2e-4                 # 13 chained clips of fp32 attention; rotations are O(1)
        assert np.array_equal(res["obj_trans"].cpu().numpy(), g["seq_obj_trans"]) and bool((res["obj_scales"] == 1).all())
    assert np.array_equal(outs[1], outs[2]) and np.abs(outs[0] - outs[1]).max() < 1e-6
    angles, _, rot6d = R.autoreg_infill(sd, opt, *seq, occ_thres=0.5)
    assert np.abs(outs[0] - angles).max() < 2e-4
    # a different threshold changes the plan's masks but not the graph
    res2 = drv.infill(*seq, occ_thres=0.2)
    ang2, _, _ = R.autoreg_infill(sd, opt, *seq, occ_thres=0.2)
    assert np.abs(res2["obj_angles"].cpu().numpy() - ang2).max() < 2e-4


def test_short_and_unseeded_sequences():
    _need_gpu()
    from vistracker_b200.infill import CondMotionInfillAutoreg
    g, opt, sd = _gold()
    net = _model(opt, sd)
    drv = CondMotionInfillAutoreg(net)
    seq = list(synthetic_infill_sequence(120, seed=2, occluded=((40, 70),)))      # shorter than one clip: a single 120-frame pass
    assert drv.clip_plan(120) == [(0, 120, 0)]
    res = drv.infill(*seq)
    ang, _, _ = R.autoreg_infill(sd, opt, *seq)
    assert np.abs(res["obj_angles"].cpu().numpy() - ang).max() < 1e-4
    assert drv.clip_plan(400)[-1] == (240, 160, 30) and len(drv.clip_plan(400)) == 10
    seq[4] = np.full(120, 0.1, np.float32)
    assert drv.infill(*seq) is None


def test_rejections():
    _need_gpu()
    from vistracker_b200 import _lib
    from vistracker_b200.infill import ConditionalMInfiller
    g, opt, sd = _gold()
    with pytest.raises(RuntimeError, match="load_state_dict"):
        ConditionalMInfiller(opt, device="cuda:0")(g["b_data_smpl"], None, g["b_data_obj"], None)
    bad = dict(sd); bad.pop("predictor.2.bias")
    with pytest.raises(RuntimeError, match="missing"):
        ConditionalMInfiller(opt, device="cuda:0").load_state_dict(bad)
    net = _model(opt, sd)
    with pytest.raises(RuntimeError, match="expected"):
        net(g["b_data_obj"], None, g["b_data_obj"], None)
    x = torch.zeros(300, 96, device="cuda")
    with pytest.raises(RuntimeError, match="at most 256"):
        _lib.call("vt_infill_attn", _lib.ptr(x), None, 1, 300, 32, 2, _lib.ptr(x), _lib.stream_ptr())
