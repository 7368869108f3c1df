"""Where does a sampling job's wall time go? (resident vs e2e; sync + perf_counter around phases)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "afford-motion_b200")); sys.path.insert(0, ROOT)
import torch
import bench as Bn
from amb200 import synth
from amb200.config import cmdm_model_cfg, full_cfg
from models.base import create_model_and_diffusion
from models.functions import set_text_feature_provider
dev = torch.device("cuda:0")
nd = int(sys.argv[1]) if len(sys.argv) > 1 else 200
model, diff = create_model_and_diffusion(full_cfg(cmdm_model_cfg(Bn.NPTS), steps=nd), device=dev)
model.load_state_dict(synth.fill_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=0), strict=False)
model.to(dev).eval()
host = Bn.synth_host_inputs(0)
txt = host["text"].to(dev)
set_text_feature_provider(lambda raw: txt)
kw = dict(c_text=host["texts"], c_pc_xyz=host["xyz"].to(dev), c_pc_contact=host["contact"].to(dev), x_mask=host["x_mask"].to(dev))
def sync(): torch.cuda.synchronize()
for rep in range(3):
    model._cond_cache = None
    sync(); t0 = time.perf_counter()
    cond = model.encode_condition(Bn.T, **kw); sync(); t1 = time.perf_counter()
    out = diff.p_sample_loop(model, (Bn.B, Bn.T, Bn.DM), clip_denoised=False, model_kwargs=kw); sync(); t2 = time.perf_counter()
    cpu = out.to("cpu"); t3 = time.perf_counter()
    print(f"rep {rep}: encode_condition {1e3*(t1-t0):.1f} ms | p_sample_loop({nd} steps) {1e3*(t2-t1):.1f} ms ({1e3*(t2-t1)/nd:.3f} ms/step) | D2H {1e3*(t3-t2):.1f} ms")
# python-side overhead of the graph loop: time replays only
import torch
g_times = []
