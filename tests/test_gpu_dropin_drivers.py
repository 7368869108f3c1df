"""The drop-in claim, executed: the REFERENCE's own drivers — `test.py::test`
(test.py:14-130) and `utils/training.py::TrainLoop.run_loop` (utils/training.py:118-180), unmodified files staged under oracle/_ref
— run against the drop-in `models` / `diffusion` packages of afford-motion_b200/ on the GPU.  Stand-ins only where the reference
needs things that do not exist offline: hydra / omegaconf / natsort (import stubs), `datasets.base` (a synthetic dataset with the
reference's batch dict schema), `utils.evaluate` (a recording evaluator), CLIP (feature provider)."""
import os
import types

import pytest
import torch

from amb200 import synth
from amb200.config import AttrDict, cmdm_model_cfg, diffusion_cfg
from oracle import ref_runtime

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
N, T, DM, K = 1024, 196, 263, 2


class _SynthDataset(torch.utils.data.Dataset):
    """Batch dict schema of datasets/humanml3d.py:740-801 (CMDM/H3D): x, x_mask, c_pc_xyz, c_pc_contact ([k,N,6] in test, [N,6] in train),
    c_text, info_*."""

    def __init__(self, n, phase):
        self.n, self.phase = n, phase

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        contact = synth.contact_map(K, N, seed=100 + i)
        return {"x": synth.motion_noise(1, T, DM, seed=i)[0], "x_mask": synth.motion_mask(2, T, seed=i)[1],
                "c_pc_xyz": synth.scene_points(1, N, seed=i)[0], "c_pc_contact": contact if self.phase == "test" else contact[0],
                "c_text": f"prompt {i}", "info_index": i}

    def get_dataloader(self, **kw):
        return torch.utils.data.DataLoader(self, **kw)


class _Evaluator:
    def __init__(self):
        self.k_samples, self.num_k_samples, self.eval_nbatch = K, 2, 2
        self.got = None

    def evaluate(self, sample_list, k_samples_list, save_dir, dataloader, device=None):
        self.got = (sample_list, k_samples_list)

    def report(self, save_dir):
        pass


@pytest.fixture(scope="module")
def drivers():
    if not ref_runtime.available():
        pytest.skip("oracle/_ref is not staged (python oracle/build_ref.py needs /root/reference)")
    ev = _Evaluator()
    dsb = types.ModuleType("datasets.base")
    dsb.create_dataset = lambda cfg, phase, gpu=None, **kw: _SynthDataset(4, phase)
    evm = types.ModuleType("utils.evaluate")
    evm.create_evaluator = lambda task_cfg, device=None: ev
    rtrain, rtest = ref_runtime.load_reference_drivers(dsb, evm)
    import models.base as mb
    assert "afford-motion_b200" in mb.__file__, "the drivers must import the DROP-IN models package"
    return rtrain, rtest, ev


def _cfg(tmp, steps):
    return AttrDict(dict(model=cmdm_model_cfg(N), diffusion=diffusion_cfg(steps), gpu=0, seed=2023, eval_dir=os.path.join(tmp, "eval"),
                         exp_dir=tmp, log_dir=os.path.join(tmp, "log"), ckpt_dir=os.path.join(tmp, "ckpt"),
                         task=dict(dataset=dict(name="synthetic"), test=dict(batch_size=2, num_workers=0),
                                   train=dict(lr=1e-4, max_steps=3, log_every_step=1, save_every_step=2, resume_ckpt=None,
                                              weight_decay=0.0, lr_anneal_steps=0))))


def test_reference_test_py_drives_the_dropin(drivers, tmp_path):
    from models.base import create_model_and_diffusion
    from models.functions import set_text_feature_provider
    rtrain, rtest, ev = drivers
    cfg = _cfg(str(tmp_path), steps=12)
    os.makedirs(cfg.ckpt_dir, exist_ok=True)
    m, _ = create_model_and_diffusion(cfg, device=DEV)
    sd = synth.fill_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=0)
    torch.save({k: v for k, v in sd.items() if "text_model" not in k}, os.path.join(cfg.ckpt_dir, "model000100.pt"))  # training.py:92-99 format
    txt = synth.text_features(8, seed=61)
    set_text_feature_provider(lambda raw: txt[: len(raw)])
    try:
        rtest.main(cfg)  # test.py:14-130 through its own main() (test.py:105-129; hydra.main is a no-op decorator here), unmodified
    finally:
        set_text_feature_provider(None)
    samples, ksamples = ev.got
    assert len(samples) == 4 and len(ksamples) == 2          # 2 batches of 2; the first batch is the k-sample batch
    for r in samples:
        assert r["sample"].shape == (T, DM) and bool((r["sample"] == r["sample"]).all())
        assert r["c_pc_contact"].shape == (K, N, 6) and isinstance(r["c_text"], str)
    assert ksamples[0]["k_samples"].shape == (K, T, DM)
    # the k-th repeat used the k-th contact map (test.py:92): different conditioning -> different samples
    assert abs(ksamples[0]["k_samples"][0] - ksamples[0]["k_samples"][1]).max() > 1e-4


def test_reference_trainloop_drives_the_dropin(drivers, tmp_path):
    from models.base import create_model_and_diffusion
    from models.functions import set_text_feature_provider
    import utils.io as rio
    rtrain, rtest, ev = drivers
    cfg = _cfg(str(tmp_path), steps=1000)
    os.makedirs(cfg.ckpt_dir, exist_ok=True)
    model, diffusion = create_model_and_diffusion(cfg, device=DEV)
    model.load_state_dict(synth.fill_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=0), strict=False)
    model.to(DEV)
    from datasets.misc import collate_fn_general  # the reference's collate (datasets/misc.py:5-13)
    loader = _SynthDataset(4, "train").get_dataloader(batch_size=2, collate_fn=collate_fn_general, shuffle=False)
    written = []
    board = rio.Board()
    board.board = types.SimpleNamespace(write=lambda d: written.append(dict(d)), close=lambda: None)
    txt = synth.text_features(8, seed=62)
    set_text_feature_provider(lambda raw: txt[: len(raw)])
    before = {k: v.detach().clone() for k, v in model.named_parameters() if k in ("motion_layer.weight", "contact_encoder.enc1.0.linear.weight")}
    try:
        loop = rtrain.TrainLoop(cfg=cfg.task.train, model=model, diffusion=diffusion, dataloader=loader, device=DEV, save_dir=cfg.ckpt_dir,
                                gpu=0, is_distributed=False)
        assert isinstance(loop.optimizer, torch.optim.AdamW)  # utils/training.py:48-50, unmodified
        loop.run_loop()
    finally:
        set_text_feature_provider(None)
    assert loop.step == 4 and len(written) == 3 and all(w["train/loss"] == w["train/loss"] for w in written)
    for k, v in before.items():
        assert (dict(model.named_parameters())[k].detach() - v).abs().max() > 0, f"{k} was not updated"
    saved = torch.load(os.path.join(cfg.ckpt_dir, "model000002.pt"))  # _save at step 2 (training.py:92-108)
    assert "motion_layer.weight" in saved and not any("text_model" in k for k in saved)
    assert os.path.exists(os.path.join(cfg.ckpt_dir, "opt.pt"))
