import sys, os
sys.argv=[sys.argv[0]]
ROOT="/root/repo"; sys.path.insert(0, ROOT); sys.path.insert(0, ROOT+"/tools")
import numpy as np, torch
from _inputs import load_assets
from vistracker_b200 import CHORETriplaneVisibility, default_options, resolve_dims
from vistracker_b200.recon_fit import SMPLParams
from vistracker_b200.smpl import LandmarkRegressor, SMPL_Layer
from vistracker_b200.synth import synthetic_recon_batch, synthetic_state_dict
from vistracker_b200.synth_smpl import synthetic_smplh, synthetic_body_mesh
dev=torch.device("cuda",0); B=96
a, reg = load_assets()
net = CHORETriplaneVisibility(default_options(), device=dev).eval()
model = synthetic_smplh(seed=3)
layer = SMPL_Layer.from_buffers(model, model["parents"], dev)
body25 = LandmarkRegressor(np.stack([reg[0], reg[1]]), reg[2], reg[3], dev)
h = synthetic_recon_batch(B, seed=4)
smpl = SMPLParams(layer, body25, h["pose"], h["betas"], h["trans"])
with torch.no_grad(): verts = smpl()[0]
bc=h["body_center"].to(dev); cc=h["crop_center"].to(dev)
cam=net._cam7.cpu().numpy() if hasattr(net,'_cam7') and torch.is_tensor(net._cam7) else None
print("cam7", net._cam7 if cam is None else cam)
def stats(v,tag):
    c=v-bc[:,None]
    inside=(c.abs()<=1).all(-1)
    print(tag,"tri-inside frac",inside.float().mean().item(),"extent", c.abs().amax((0,1)).tolist(), "std", c.std((0,1)).tolist())
    # warps of 8 consecutive points
    N=v.shape[1]; n8=(N//8)*8
    w=inside[:,:n8].reshape(B,-1,8).all(-1).float().mean().item()
    print(tag,"8-point groups all inside", w)
stats(verts,"smpl")
bv,_=synthetic_body_mesh()
v2=(torch.from_numpy(bv)[None]*0.9+h["body_center"][:,None]).to(dev)
stats(v2,"prof_tool")
