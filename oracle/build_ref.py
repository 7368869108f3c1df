"""ORACLE (test infrastructure): recipe that stages the reference's own Python modules for the hot path under oracle/_ref/.

The reference is a plain Python tree (no setup.py / pyproject), so "building" it is a file copy: this script copies the
UNMODIFIED files listed below from /root/reference into oracle/_ref/ (git-ignored, NOT gpurun-ignored: it travels to the GPU
box like a built .so, /root/reference does not exist there).  Nothing is copied into tracked paths, and nothing under
afford-motion_b200/ ever imports from here; users are `bench.py --impl reference` / the `cpu_baseline` leg (the reference's
own CMDM / CDM modules timed on the host cores) and the drop-in driver tests (tests/test_gpu_dropin_drivers.py: the reference's
utils/training.py::TrainLoop and test.py driving the drop-in models).  oracle/ref_runtime.py provides the import stubs for the
reference's unavailable third-party imports (omegaconf, hydra, clip, pointops_cuda, smplkit, natsort).

    python oracle/build_ref.py        # idempotent; prints what it staged
"""
import filecmp
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
OUT = os.path.join(HERE, "_ref")
FILES = [
    "models/__init__.py", "models/base.py", "models/cdm.py", "models/cmdm.py", "models/modules.py", "models/functions.py",
    "models/scene_models/pointops.py", "models/scene_models/pointtransformer.py",
    "diffusion/gaussian_diffusion.py", "diffusion/respace.py", "diffusion/resample.py", "diffusion/nn.py", "diffusion/losses.py",
    "utils/registry.py", "utils/misc.py", "utils/training.py", "utils/io.py",
    "datasets/misc.py",
    "test.py", "train_ddp.py",
]


def available() -> bool:
    return all(os.path.exists(os.path.join(OUT, f)) for f in FILES) and os.path.exists(os.path.join(OUT, "diffusion", "__init__.py"))


def build(verbose: bool = False) -> bool:
    """Stage the files when /root/reference is present (the build container); on the GPU box just report what is there."""
    if not os.path.isdir(REF):
        return available()
    n = 0
    for f in FILES:
        src, dst = os.path.join(REF, f), os.path.join(OUT, f)
        if not os.path.exists(src):
            raise FileNotFoundError(src)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not (os.path.exists(dst) and filecmp.cmp(src, dst, shallow=False)):
            shutil.copyfile(src, dst)
            n += 1
    # `diffusion/` is a namespace package in the reference (no __init__.py); a REGULAR package of the same name anywhere on sys.path
    # (the drop-in afford-motion_b200/diffusion) would shadow it, so the staged copy gets an empty marker file.
    marker = os.path.join(OUT, "diffusion", "__init__.py")
    if not os.path.exists(marker):
        open(marker, "w").close()
    if verbose:
        print(f"oracle/_ref: {len(FILES)} reference files staged ({n} copied or refreshed)")
    return True


if __name__ == "__main__":
    build(verbose=True)
