"""Full-length CPU run of the headline workload with the REFERENCE's own modules ("run it fully at least once"):
CMDM 1000-step DDPM sampling, batch 32, T=196, D=263, N=8192 through the reference's `diffusion.p_sample_loop(model, ...)`
(oracle/_ref staged by oracle/build_ref.py; pointops_cuda served by oracle/pointops_ref.c, CLIP by a synthetic feature provider).
Conditioning is recomputed on every step exactly as models/cmdm.py:133-149 does.  Prints one JSON line.
    python tools/cpu_full_run.py [--steps 1000] [--threads N] > profiles/r2_cpu_full_1000step_reference.json"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "afford-motion_b200")]
import torch
ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=1000)
ap.add_argument("--threads", type=int, default=os.cpu_count())
ap.add_argument("--batch", type=int, default=32)
a = ap.parse_args()
torch.set_num_threads(a.threads)
os.environ["OMP_NUM_THREADS"] = str(min(a.threads, a.batch))
from amb200 import synth
from amb200.config import cmdm_model_cfg
from oracle import ref_runtime
B, T, DM, N = a.batch, 196, 263, 8192
txt = synth.text_features(B, seed=2023)
rbase, rgd = ref_runtime.reference_models(lambda raw: txt[: len(raw)])
model, diff = rbase.create_model_and_diffusion(ref_runtime.full_cfg(cmdm_model_cfg(N), steps=a.steps), device="cpu")
model.load_state_dict(synth.fill_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=0), strict=False)
model.eval()
kw = dict(c_text=[f"p{i}" for i in range(B)], c_pc_xyz=synth.scene_points(B, N, seed=2023), c_pc_contact=synth.contact_map(B, N, seed=2023),
          x_mask=synth.motion_mask(B, T, seed=2023, all_valid=True))
torch.manual_seed(2023)
t0 = time.perf_counter()
out = diff.p_sample_loop(model, (B, T, DM), clip_denoised=False, noise=None, model_kwargs=kw, progress=False)
dt = time.perf_counter() - t0
print(json.dumps({"workload": f"CMDM {a.steps}-step DDPM sampling, batch={B}, T=196, D=263, N=8192, reference modules on CPU (conditioning recomputed per step)",
                  "seconds": dt, "denoise_steps_per_s": a.steps / dt, "motions_per_s": B / dt, "threads": a.threads, "host_cpus": os.cpu_count(),
                  "finite": bool(torch.isfinite(out).all()), "kind": "reference"}))
