"""Host-side logic that needs no GPU: mini-batch combination and key-point scaling of the fit_recon driver, the HVOP-Net clip plan / masks,
the 8-bit resize tables of the frame preparation."""
import numpy as np
import torch

from oracle import frameio_ref as FR
from oracle import infill_ref as IR


def test_combine_mini_batches_and_kpt_scaling():
    from vistracker_b200.recon_driver import combine_mini_batches, scale_body_kpts
    mk = lambda b, n: {t: {"points": torch.rand(b, n, 3), "parts": torch.randint(0, 14, (b, n)), "centers": torch.rand(b, 6),
                            "pca_axis": torch.rand(b, 3, 3), "visibility": torch.rand(b, 1)} for t in ("human", "object")}
    a, b = mk(2, 50), mk(1, 40)
    c = combine_mini_batches([a, b], 40)                                            # recon_fit_behave.py:152-183
    assert c["human"]["points"].shape == (3, 40, 3) and c["object"]["parts"].shape == (3, 40) and c["object"]["centers"].shape == (3, 6)
    assert torch.equal(c["human"]["points"][:2], a["human"]["points"][:, :40]) and torch.equal(c["object"]["pca_axis"][2:], b["object"]["pca_axis"])
    k = torch.tensor([[[1024.0, 768.0, 0.5], [424.0, 168.0, 1.0]]])
    out = scale_body_kpts(k, torch.tensor([[1024.0, 768.0]]))                       # recon_fit_base.py:397-409 with unit scales
    assert torch.allclose(out[0], torch.tensor([[256.0, 256.0, 0.5], [0.0, 0.0, 1.0]]))
    out2 = scale_body_kpts(k, torch.tensor([[1024.0, 768.0]]), resize_scale=torch.tensor([2.0]), crop_scale=torch.tensor([2.0]))
    assert torch.allclose(out2[0, 0, :2], torch.tensor([(2048.0 - 1024.0 + 1200.0) * 512 / 2400, (1536.0 - 768.0 + 1200.0) * 512 / 2400]))


def test_infill_clip_plan_and_masks_follow_the_reference_loop():
    from vistracker_b200.infill import CondMotionInfillAutoreg
    drv = CondMotionInfillAutoreg.__new__(CondMotionInfillAutoreg)
    drv.clip_len, drv.window, drv.init_thres = 180, 30, 0.5
    for L in (100, 180, 200, 400, 1500):
        plan = drv.clip_plan(L)
        ref = [(0, min(180, L), 0)] + [(i, min(180, L - i), min(30, L - i)) for i in range(0, L - 180 + 1 + 30, 30)]   # test_infill_autoreg.py:93,116-117
        assert plan == ref, L
        assert all(s + T <= L and T > 0 for s, T, _ in plan)
    occ = np.linspace(0, 1, 400).astype(np.float32)
    rows = drv._masks(occ, 0.3, drv.clip_plan(400))
    assert rows.shape == (len(drv.clip_plan(400)), 180)
    assert np.array_equal(rows[0].astype(bool), occ[:180] < 0.5)                     # first clip: init_thres, no context frames
    m2 = occ[30:210] < 0.3; m2[:30] = False
    assert np.array_equal(rows[2].astype(bool), m2)
    assert rows[-1, 160:].sum() == 0                                                 # the short last clip is padded with "visible"


def test_resize_tables_match_the_oracle_tables():
    from vistracker_b200.frameio import resize_table
    for d, s in ((512, 1200), (64, 150), (100, 100), (33, 200), (5, 3)):
        t = resize_table(d, s)
        i, a0, a1 = FR._coeffs(d, s)
        assert t.dtype == np.int32 and np.array_equal(t[0], i) and np.array_equal(t[1], a0) and np.array_equal(t[2], a1)
        assert int(t[0].max()) <= s - 1 and bool(((t[1] + t[2]) >= 2047).all()) and bool(((t[1] + t[2]) <= 2049).all())


def test_position_embedding_host_table_equals_the_oracle():
    from vistracker_b200.infill import position_embedding
    for L, D in ((180, 128), (180, 32), (160, 160), (47, 33), (1, 8)):
        assert torch.equal(position_embedding(L, D), IR.position_embedding(L, D))
