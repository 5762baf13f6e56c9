"""Developer diagnostic (run on the GPU box: `python tests/diag_loops.py`): both optimisation loops, CUDA-graph and PyTorch-glue execution,
against the reference-loop goldens -- prints per-step deviations instead of asserting."""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))
from recon_problem import B, make_loop_extras, make_problem  # noqa: E402
from vistracker_b200 import CHORETriplaneVisibility, default_options, resolve_dims  # noqa: E402
from vistracker_b200.recon_fit import Priors, ReconFitterTriVisFull, SMPLParams  # noqa: E402
from vistracker_b200.smpl import LandmarkRegressor, SMPL_Layer  # noqa: E402
from vistracker_b200.synth import synthetic_state_dict  # noqa: E402

np.set_printoptions(linewidth=220, precision=3, suppress=False)
d = make_problem(); e = make_loop_extras(d)
dims = resolve_dims(default_options())
net = CHORETriplaneVisibility(default_options(), device="cuda:0").eval()
net.load_state_dict(synthetic_state_dict(dims, seed=0))
net.filter(d["images"].cuda())
layer = SMPL_Layer.from_buffers(d["model"], d["model"]["parents"], "cuda:0")
reg = LandmarkRegressor(np.stack([d["reg"][0], d["reg"][1]]), d["reg"][2], d["reg"][3], "cuda:0")
fitter = ReconFitterTriVisFull(net, Priors(d["assets"], "cuda:0"), d["labels"], scan=(e["temp_v"], e["temp_f"]))
make_smpl = lambda: SMPLParams(layer, reg, d["pose"], d["betas"], d["trans"])
c = lambda t: t.cuda()
G = lambda n: dict(np.load(os.path.join(HERE, "golden", n)))
g = G("recon_loop.npz")
for tag, kw in (("a", dict(steps_per_iter=3, max_iter=2)), ("b", dict(steps_per_iter=2, max_iter=12))):
    for mode in ("eager", "graph"):
        dd = {"part_labels": c(d["labels"])[None].repeat(B, 1), "query_dict": {"crop_center": c(d["crop"]), "body_center": c(d["body_center"])},
              "pose_init": c(d["pose_init"]), "body_kpts": c(d["body_kpts"])}
        smpl = make_smpl()
        t0 = time.perf_counter()
        try:
            smpl, scale = fitter.optimize_smpl(smpl, dd, 1, 1, 1, loop_mode=mode, **kw)
        except Exception as ex:          # noqa: BLE001
            print(f"[smpl {tag} {mode}] FAILED: {type(ex).__name__}: {ex}")
            continue
        torch.cuda.synchronize()
        h, r = np.asarray(fitter.last_hist), g[f"{tag}_hist"]
        n = min(len(h), len(r))
        print(f"[smpl {tag} {mode}] steps {len(h)} (ref {len(r)}) stopped {fitter.last_stopped} in {time.perf_counter() - t0:.2f}s")
        print("   total rel dev:", np.abs(h[:n] - r[:n]) / np.abs(r[:n]))
        T, RT = np.asarray(fitter.last_terms)[:n], g[f"{tag}_terms"][:n]
        with np.errstate(invalid="ignore"):
            print("   worst term rel dev per term", dict(zip(g["term_names"], np.nanmax(np.abs(T - RT) / np.maximum(np.abs(RT), 1e-9), 0))))
        print("   nan pattern equal:", np.array_equal(np.isnan(T), np.isnan(RT)))
        pose = torch.cat([smpl.global_pose, smpl.body_pose, smpl.hand_pose], 1).detach().cpu().numpy()
        betas = torch.cat([smpl.top_betas, smpl.other_betas], 1).detach().cpu().numpy()
        rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
        print("   final pose/betas/trans/scale rel:", rel(pose, g[f"{tag}_pose"]), rel(betas, g[f"{tag}_betas"]),
              rel(smpl.trans.detach().cpu().numpy(), g[f"{tag}_trans"]), rel(scale.cpu().numpy(), g[f"{tag}_scale"]))

g = G("recon_obj_loop.npz")
for mode in ("eager", "graph"):
    dd = {"images": c(e["images_sil"]), "query_dict": {"crop_center": c(d["crop"]), "body_center": c(d["body_center"])}, "camera_params": {},
          "crop_size": 1200, "net_input_size": d["images"].shape[-1], "smpl": make_smpl(), "obj_R": c(d["obj_R"]).requires_grad_(True),
          "obj_t": c(d["obj_t"]).requires_grad_(True), "obj_s": c(d["obj_s"]), "objects": c(d["objects"]), "occ_ratios": c(d["occ"])}
    draws = [0]

    def noise_fn():
        draws[0] += 1
        return e["noise_seq"][draws[0] - 1].cuda()
    fitter.get_opt_iters = staticmethod(lambda: {"sil": 2, "object": 2})
    t0 = time.perf_counter()
    try:
        _, R_out, t_out = fitter.optimize_smpl_object(net, dd, joint_iter=1, steps_per_iter=1, noise_fn=noise_fn, loop_mode=mode)
    except Exception as ex:          # noqa: BLE001
        import traceback; traceback.print_exc()
        print(f"[obj {mode}] FAILED: {type(ex).__name__}: {ex}")
        continue
    torch.cuda.synchronize()
    h, r = np.asarray(fitter.last_hist), g["hist"]
    n = min(len(h), len(r))
    print(f"[obj {mode}] steps {len(h)} (ref {len(r)}) draws {draws[0]} (ref {int(g['n_draws'])}) stopped {fitter.last_stopped} in {time.perf_counter() - t0:.2f}s")
    dev = np.abs(h[:n] - r[:n]) / np.abs(r[:n])
    print("   total rel dev first 8:", dev[:8], "max", dev.max(), "at", int(dev.argmax()))
    names = list(g["term_names"])
    T = np.asarray(fitter.last_terms)[:n]
    for k, name in enumerate(("otemp", "ovtemp", "mask", "scale", "trans", "object", "contact")):
        RT = g["terms"][:n, names.index(name)]
        with np.errstate(invalid="ignore", divide="ignore"):
            dv = np.abs(T[:, k] - RT) / np.maximum(np.abs(RT), 1e-9)
        print(f"   {name}: nan-pattern equal {np.array_equal(np.isnan(T[:, k]), np.isnan(RT))}, worst rel dev {np.nanmax(dv) if np.isfinite(dv).any() else 'n/a'}; first ours {T[:5, k]} ref {RT[:5]}")
    sil = dd["silhouette"]
    print("   keep/ref masks equal:", np.array_equal(sil.keep_mask.cpu().numpy(), g["keep_mask"]), np.array_equal(sil.image_ref.cpu().numpy(), g["image_ref"]),
          "K4", sil.renderer.K4[0].cpu().numpy(), "ref K", g["K_roi"][0].ravel())
    rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
    print("   df_obj_h / df_hum_o rel:", rel(dd["df_obj_h"].cpu().numpy(), g["df_obj_h"]), rel(dd["df_hum_o"].cpu().numpy(), g["df_hum_o"]),
          "parts argmax equal:", float((dd["parts_obj"].argmax(1).cpu().numpy() == g["parts_obj"].argmax(1)).mean()))
    print("   final obj_R / obj_t / rot rel:", rel(R_out.detach().cpu().numpy(), g["obj_R"]), rel(t_out.detach().cpu().numpy(), g["obj_t"]),
          rel(fitter.final_rotation(R_out).cpu().numpy(), g["rot_final"]))
