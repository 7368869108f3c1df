"""Pin the oracle's TRAINING-mode restatement (batch-statistics BatchNorm, dropout 0) against a training step executed by
the reference's own modules (tests/golden/cmdm_train_b2_n1024.npz): per-sample loss and parameter gradients."""
import json
import os

import numpy as np
import torch

from amb200 import synth
from oracle import cmdm_ref, diffusion_ref as D
from oracle.nn_ref import positional_table


def oracle_train_step(golden_dir, device="cpu"):
    shapes = json.load(open(os.path.join(golden_dir, "state_keys.json")))["CMDM"]
    sd = synth.fill_state_dict(shapes, seed=0)
    sd["timestep_embedder.pe"] = positional_table(1000, 512).unsqueeze(1)
    sd["positional_encoder.pe"] = positional_table(5000, 512).unsqueeze(1)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and not k.endswith((".pe", "running_mean", "running_var"))}
    full = dict(sd)
    full.update(params)
    B, T, Dm, N = 2, 196, 263, 1024
    xyz = synth.scene_points(B, N, seed=31, dup_frac=0.05)
    contact = synth.contact_map(B, N, seed=31)
    x0 = synth.motion_noise(B, T, Dm, seed=31)
    x_mask = synth.motion_mask(B, T, seed=31)
    x_mask[1, 100:] = True
    txt = synth.text_features(B, seed=31)
    noise = synth.step_noise((B, T, Dm), 77)
    t = torch.tensor([700, 23])
    tab = D.make_tables(D.respaced(D.cosine_betas(1000), range(1000))[0])
    x_t = D.q_sample(tab, x0, t, noise)
    pred = cmdm_ref.cmdm_forward(full, x_t, t, txt, xyz, contact, x_mask, train=True)
    loss = D.masked_mse(x0, pred, x_mask)
    loss.mean().backward()
    return loss.detach(), {k: p.grad for k, p in params.items() if p.grad is not None}, dict(xyz=xyz, contact=contact, x0=x0, x_mask=x_mask, txt=txt,
                                                                                          noise=noise, t=t)


def test_oracle_training_step_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "cmdm_train_b2_n1024.npz"))
    loss, grads, _ = oracle_train_step(golden_dir)
    np.testing.assert_allclose(loss.numpy(), g["loss"], rtol=2e-5)
    names = [str(n) for n in g["grad_names"]]
    assert set(names) == set(grads)
    mine = np.array([float(grads[n].norm()) for n in names])
    np.testing.assert_allclose(mine, g["grad_norms"], rtol=2e-3, atol=1e-7)
    for k in g.files:
        if k.startswith("grad::"):
            ref = g[k]
            err = np.abs(grads[k[6:]].numpy() - ref).max()
            assert err <= 2e-3 * max(np.abs(ref).max(), 1e-6) + 1e-7, (k, err)


def oracle_cdm_train_step(golden_dir):
    from oracle import cdm_ref
    shapes = json.load(open(os.path.join(golden_dir, "state_keys.json")))["CDM"]
    sd = synth.fill_state_dict(shapes, seed=0)
    sd["timestep_embedder.pe"] = positional_table(1000, 128).unsqueeze(1)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and not k.endswith(".pe")}
    full = dict(sd)
    full.update(params)
    B, N = 2, 1024
    xyz = synth.scene_points(B, N, seed=41)
    x0 = torch.randn(B, N, 6, generator=torch.Generator().manual_seed(41))
    noise = synth.step_noise((B, N, 6), 78)
    txt = synth.text_features(B, seed=41)
    t = torch.tensor([400, 7])
    tab = D.make_tables(D.respaced(D.cosine_betas(500), range(500))[0])
    x_t = D.q_sample(tab, x0, t, noise)
    pred = cdm_ref.cdm_forward(full, x_t, t, txt, xyz)
    loss = D.masked_mse(x0, pred, torch.zeros(B, N, dtype=torch.bool))
    loss.mean().backward()
    return loss.detach(), {k: p.grad for k, p in params.items() if p.grad is not None}, dict(xyz=xyz, x0=x0, noise=noise, txt=txt, t=t)


def test_oracle_cdm_training_step_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "cdm_train_b2_n1024.npz"))
    loss, grads, _ = oracle_cdm_train_step(golden_dir)
    np.testing.assert_allclose(loss.numpy(), g["loss"], rtol=2e-5)
    names = [str(n) for n in g["grad_names"]]
    assert set(names) == set(grads)
    np.testing.assert_allclose(np.array([float(grads[n].norm()) for n in names]), g["grad_norms"], rtol=2e-3, atol=1e-8)
    for k in g.files:
        if k.startswith("grad::"):
            ref = g[k]
            assert np.abs(grads[k[6:]].numpy() - ref).max() <= 2e-3 * max(np.abs(ref).max(), 1e-6) + 1e-7, k
