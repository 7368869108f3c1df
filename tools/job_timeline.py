"""Host-side timeline of consecutive CMDM sampling jobs (bench.py's resident job): where the host is at each phase, and the
device time per job (CUDA events).  Used to explain job-to-job variance (first job after a sync, graph capture cost)."""
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
nd = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
model, diff = create_model_and_diffusion(full_cfg(cmdm_model_cfg(Bn.NPTS), steps=nd), device=dev)
model.load_state_dict(synth.fill_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=0), strict=False)
model.to(dev).eval()
host = Bn.synth_host_inputs(0)
txt = host["text"].to(dev)
set_text_feature_provider(lambda raw: txt)
kw = dict(c_text=host["texts"], c_pc_xyz=host["xyz"].to(dev), c_pc_contact=host["contact"].to(dev), x_mask=host["x_mask"].to(dev))
def job():
    model._cond_cache = None
    return diff.p_sample_loop(model, (Bn.B, Bn.T, Bn.DM), clip_denoised=False, model_kwargs=kw)
for _ in range(2):
    job()
torch.cuda.synchronize()
for phase in range(2):
    K = 4
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    traces = []
    torch.cuda.synchronize()
    t00 = time.perf_counter()
    ev[0].record()
    for i in range(K):
        diff.trace = []
        traces.append(diff.trace)
        diff.trace.append(("job_call", time.perf_counter()))
        job()
        diff.trace.append(("job_return", time.perf_counter()))
        ev[i + 1].record()
    torch.cuda.synchronize()
    t_end = time.perf_counter()
    print(f"phase {phase}: wall {1e3*(t_end-t00):.1f} ms; device ms per job: {[round(ev[i].elapsed_time(ev[i+1]),1) for i in range(K)]}")
    for i, tr in enumerate(traces):
        print(f"  job {i}: " + " | ".join(f"{lab} +{1e3*(t-t00):.1f}" for lab, t in tr))
    diff.trace = None
    time.sleep(1.0)
