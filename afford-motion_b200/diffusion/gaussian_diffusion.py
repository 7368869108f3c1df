"""Gaussian diffusion (DDPM / DDIM) — drop-in for /root/reference/diffusion/gaussian_diffusion.py.

Same public surface (`GaussianDiffusion`, `ModelMeanType`, `ModelVarType`, `LossType`, `get_named_beta_schedule`,
`q_sample`, `p_mean_variance`, `p_sample(_loop)(_progressive)`, `ddim_sample(_loop)(_progressive)`,
`training_losses` -> {'loss','mse'}), same float64 host tables (gaussian_diffusion.py:119-170).

What is different (B200-first):
  * the fp32 coefficient tables live on the device once (the reference copies 2-4 float64 tables H2D every step,
    `_extract_into_tensor` :829-842); the whole per-step update is ONE fused kernel (am_p_sample_update /
    am_ddim_update) with in-kernel Philox noise or injected noise;
  * for models that expose `sampler_begin` (CDM / CMDM of this package) the sampling loops keep x_t and the
    timestep on the device and replay ONE captured CUDA graph per denoise step: no per-step H2D copies,
    `.item()` syncs, `th.tensor([i]*B)` uploads or text / point-cloud re-encoding.
Configured path: START_X or EPSILON mean, FIXED_SMALL / FIXED_LARGE variance, MSE loss (configs/default.yaml:31-40).
Learned-sigma / KL branches (:260-277, :710-743) are dead under every reference config and raise NotImplementedError.
"""
import enum
import math

import numpy as np
import torch as th

from amb200 import ops


def get_named_beta_schedule(schedule_name, num_diffusion_timesteps):
    """gaussian_diffusion.py:19-43."""
    if schedule_name == "linear":
        scale = 1000 / num_diffusion_timesteps
        return np.linspace(scale * 0.0001, scale * 0.02, num_diffusion_timesteps, dtype=np.float64)
    if schedule_name == "cosine":
        return betas_for_alpha_bar(num_diffusion_timesteps, lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2)
    raise NotImplementedError(f"unknown beta schedule: {schedule_name}")


def betas_for_alpha_bar(num_diffusion_timesteps, alpha_bar, max_beta=0.999):
    """gaussian_diffusion.py:46-63."""
    n = num_diffusion_timesteps
    return np.array([min(1 - alpha_bar((i + 1) / n) / alpha_bar(i / n), max_beta) for i in range(n)])


class ModelMeanType(enum.Enum):
    PREVIOUS_X = enum.auto()
    START_X = enum.auto()
    EPSILON = enum.auto()


class ModelVarType(enum.Enum):
    LEARNED = enum.auto()
    FIXED_SMALL = enum.auto()
    FIXED_LARGE = enum.auto()
    LEARNED_RANGE = enum.auto()


class LossType(enum.Enum):
    MSE = enum.auto()
    RESCALED_MSE = enum.auto()
    KL = enum.auto()
    RESCALED_KL = enum.auto()

    def is_vb(self):
        return self in (LossType.KL, LossType.RESCALED_KL)


def _draw_seed() -> int:
    """Philox seed drawn from torch's default CPU generator, so torch.manual_seed() controls sampling noise."""
    return int(th.randint(0, 2 ** 62, (1,), dtype=th.int64).item())


class GaussianDiffusion:
    def __init__(self, *, betas, model_mean_type, model_var_type, loss_type, rescale_timesteps=False):
        self.model_mean_type = model_mean_type
        self.model_var_type = model_var_type
        self.loss_type = loss_type
        self.rescale_timesteps = rescale_timesteps

        betas = np.array(betas, dtype=np.float64)
        self.betas = betas
        assert betas.ndim == 1, "betas must be 1-D"
        assert (betas > 0).all() and (betas <= 1).all()
        self.num_timesteps = int(betas.shape[0])

        alphas = 1.0 - betas
        ac = np.cumprod(alphas, axis=0)
        self.alphas_cumprod = ac
        self.alphas_cumprod_prev = np.append(1.0, ac[:-1])
        self.alphas_cumprod_next = np.append(ac[1:], 0.0)
        self.sqrt_alphas_cumprod = np.sqrt(ac)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - ac)
        self.log_one_minus_alphas_cumprod = np.log(1.0 - ac)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / ac)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / ac - 1)
        self.posterior_variance = betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - ac)
        self.posterior_log_variance_clipped = np.log(np.append(self.posterior_variance[1], self.posterior_variance[1:]))
        self.posterior_mean_coef1 = betas * np.sqrt(self.alphas_cumprod_prev) / (1.0 - ac)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(alphas) / (1.0 - ac)
        self._dev = {}
        self.sample_offset = 0   # global index of this rank's first sample (rank-count-invariant Philox streams)
        self.last_launches = 0   # kernels launched by the last device-resident loop (eager + graph replays)

    # ------------------------------------------------------------------ device tables (once per device)
    def _tables(self, device):
        key = str(device)
        tab = self._dev.get(key)
        if tab is None:
            f = lambda a: th.from_numpy(np.ascontiguousarray(a)).float().to(device)  # fp64 -> fp32 like `.float()` at :839
            if self.model_var_type == ModelVarType.FIXED_LARGE:
                var = np.append(self.posterior_variance[1], self.betas[1:])
                logvar = np.log(var)
            else:
                var, logvar = self.posterior_variance, self.posterior_log_variance_clipped
            tab = dict(coef1=f(self.posterior_mean_coef1), coef2=f(self.posterior_mean_coef2), var=f(var), logvar=f(logvar),
                       sqrt_ac=f(self.sqrt_alphas_cumprod), sqrt_1mac=f(self.sqrt_one_minus_alphas_cumprod),
                       sqrt_recip_ac=f(self.sqrt_recip_alphas_cumprod), sqrt_recipm1_ac=f(self.sqrt_recipm1_alphas_cumprod),
                       ac=f(self.alphas_cumprod), ac_prev=f(self.alphas_cumprod_prev))
            self._dev[key] = tab
        return tab

    def _check_supported(self):
        if self.model_var_type in (ModelVarType.LEARNED, ModelVarType.LEARNED_RANGE):
            raise NotImplementedError("learned-sigma diffusion is outside the reference's configured path (learn_sigma=false)")
        if self.model_mean_type == ModelMeanType.PREVIOUS_X:
            raise NotImplementedError("ModelMeanType.PREVIOUS_X is not reachable from models/base.py:40-43")

    def _scale_timesteps(self, t):
        if self.rescale_timesteps:
            return t.float() * (1000.0 / self.num_timesteps)
        return t

    @staticmethod
    def _t32(t):
        return t.to(dtype=th.int32).contiguous()

    # ------------------------------------------------------------------ forward process
    def q_sample(self, x_start, t, noise=None):
        """:189-207."""
        x_start = x_start.float().contiguous()
        if noise is None:
            noise = th.empty_like(x_start)
            ops.randn_(noise, x_start[0].numel(), x_start.shape[0], 0, _draw_seed(), 0)
        assert noise.shape == x_start.shape
        tab = self._tables(x_start.device)
        out = th.empty_like(x_start)
        return ops.q_sample(x_start, noise.float().contiguous(), out, tab["sqrt_ac"], tab["sqrt_1mac"], self._t32(t))

    def _predict_xstart_from_eps(self, x_t, t, eps):
        """:329-334 (EPSILON-mean models only)."""
        tab = self._tables(x_t.device)
        sh = (-1,) + (1,) * (x_t.dim() - 1)
        return tab["sqrt_recip_ac"][t].view(sh) * x_t - tab["sqrt_recipm1_ac"][t].view(sh) * eps

    def _predict_eps_from_xstart(self, x_t, t, pred_xstart):
        """:346-350."""
        tab = self._tables(x_t.device)
        sh = (-1,) + (1,) * (x_t.dim() - 1)
        return (tab["sqrt_recip_ac"][t].view(sh) * x_t - pred_xstart) / tab["sqrt_recipm1_ac"][t].view(sh)

    def q_posterior_mean_variance(self, x_start, x_t, t):
        """:209-231."""
        tab = self._tables(x_t.device)
        mean = th.empty_like(x_t)
        zero = th.zeros_like(x_t)
        ops.p_sample_update(x_start.float().contiguous(), x_t.float().contiguous(), mean, zero, tab["coef1"], tab["coef2"], tab["logvar"],
                            self._t32(t), 1)
        sh = (-1,) + (1,) * (x_t.dim() - 1)
        return mean, tab["var"][t].view(sh).expand(x_t.shape), tab["logvar"][t].view(sh).expand(x_t.shape)

    # ------------------------------------------------------------------ reverse process, single step API
    def _model_xstart(self, model, x, t, clip_denoised, denoised_fn, model_kwargs):
        self._check_supported()
        out = model(x, self._scale_timesteps(t), **(model_kwargs or {}))
        if self.model_mean_type == ModelMeanType.EPSILON:
            out = self._predict_xstart_from_eps(x, t, out)
        if denoised_fn is not None:
            out = denoised_fn(out)
        if clip_denoised:
            out = out.clamp(-1, 1)
        return out.float().contiguous()

    def p_mean_variance(self, model, x, t, clip_denoised=True, denoised_fn=None, model_kwargs=None):
        """:233-327 (START_X / EPSILON with fixed variance)."""
        B = x.shape[0]
        assert t.shape == (B,)
        pred_xstart = self._model_xstart(model, x, t, clip_denoised, denoised_fn, model_kwargs)
        mean, var, logvar = self.q_posterior_mean_variance(pred_xstart, x, t)
        return {"mean": mean, "variance": var, "log_variance": logvar, "pred_xstart": pred_xstart}

    def p_sample(self, model, x, t, clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None, noise=None):
        """:396-440.  `noise` (extra, optional) injects eps instead of drawing it (parity harnesses)."""
        if cond_fn is not None:
            raise NotImplementedError("classifier guidance (cond_fn) is unused by the reference drivers")
        pred_xstart = self._model_xstart(self._wrap_model(model), x, t, clip_denoised, denoised_fn, model_kwargs)
        tab = self._tables(x.device)
        sample = th.empty_like(pred_xstart)
        ops.p_sample_update(pred_xstart, x.float().contiguous(), sample, None if noise is None else noise.float().contiguous(),
                            tab["coef1"], tab["coef2"], tab["logvar"], self._t32(t), 1, seed=0 if noise is not None else _draw_seed())
        return {"sample": sample, "pred_xstart": pred_xstart}

    def ddim_sample(self, model, x, t, clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None, eta=0.0, noise=None):
        """:538-586."""
        if cond_fn is not None:
            raise NotImplementedError("classifier guidance (cond_fn) is unused by the reference drivers")
        pred_xstart = self._model_xstart(self._wrap_model(model), x, t, clip_denoised, denoised_fn, model_kwargs)
        tab = self._tables(x.device)
        sample = th.empty_like(pred_xstart)
        ops.ddim_update(pred_xstart, x.float().contiguous(), sample, None if noise is None else noise.float().contiguous(),
                        tab["sqrt_recip_ac"], tab["sqrt_recipm1_ac"], tab["ac"], tab["ac_prev"], eta, self._t32(t), 1,
                        seed=0 if noise is not None else _draw_seed())
        return {"sample": sample, "pred_xstart": pred_xstart}

    def _wrap_model(self, model):  # SpacedDiffusion overrides
        return model

    # ------------------------------------------------------------------ loops
    def p_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None,
                      device=None, progress=False):
        """:442-486."""
        final = None
        for sample in self.p_sample_loop_progressive(model, shape, noise=noise, clip_denoised=clip_denoised, denoised_fn=denoised_fn,
                                                     cond_fn=cond_fn, model_kwargs=model_kwargs, device=device, progress=progress,
                                                     _only_final=True):
            final = sample
        return final["sample"]

    def ddim_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None,
                         device=None, progress=False, eta=0.0):
        """:626-657."""
        final = None
        for sample in self.ddim_sample_loop_progressive(model, shape, noise=noise, clip_denoised=clip_denoised, denoised_fn=denoised_fn,
                                                        cond_fn=cond_fn, model_kwargs=model_kwargs, device=device, progress=progress,
                                                        eta=eta, _only_final=True):
            final = sample
        return final["sample"]

    def p_sample_loop_progressive(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None,
                                  device=None, progress=False, _only_final=False):
        """:488-536."""
        yield from self._loop("ddpm", model, shape, noise, clip_denoised, denoised_fn, cond_fn, model_kwargs, device, progress, 0.0,
                              _only_final)

    def ddim_sample_loop_progressive(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                                     model_kwargs=None, device=None, progress=False, eta=0.0, _only_final=False):
        """:659-708."""
        yield from self._loop("ddim", model, shape, noise, clip_denoised, denoised_fn, cond_fn, model_kwargs, device, progress, eta,
                              _only_final)

    def _loop(self, kind, model, shape, noise, clip_denoised, denoised_fn, cond_fn, model_kwargs, device, progress, eta, only_final):
        self._check_supported()
        if device is None:
            device = next(model.parameters()).device
        device = th.device(device)
        assert isinstance(shape, (tuple, list))
        shape = tuple(shape)
        B = shape[0]
        seed = _draw_seed()
        if noise is not None:
            img = noise.to(device).float().contiguous().clone()
        else:
            img = th.empty(*shape, device=device)
            ops.randn_(img, img[0].numel(), B, self.sample_offset, seed, 0xFFFFFFFF)  # x_T (subsequence distinct from every step's)
        indices = list(range(self.num_timesteps))[::-1]
        fast = (hasattr(model, "sampler_begin") and cond_fn is None and denoised_fn is None and not clip_denoised
                and self.model_mean_type == ModelMeanType.START_X and not self.rescale_timesteps)
        if fast:
            yield from self._fast_loop(kind, model, img, model_kwargs or {}, eta, seed, progress, only_final)
            return
        if progress:
            from tqdm.auto import tqdm
            indices = tqdm(indices)
        for i in indices:
            t = th.full((B,), i, device=device, dtype=th.long)
            with th.no_grad():
                if kind == "ddpm":
                    out = self.p_sample(model, img, t, clip_denoised=clip_denoised, denoised_fn=denoised_fn, cond_fn=cond_fn,
                                        model_kwargs=model_kwargs)
                else:
                    out = self.ddim_sample(model, img, t, clip_denoised=clip_denoised, denoised_fn=denoised_fn, cond_fn=cond_fn,
                                           model_kwargs=model_kwargs, eta=eta)
                yield out
                img = out["sample"]

    # ------------------------------------------------------------------ device-resident loop (CUDA graph per step)
    def _fast_loop(self, kind, model, img, model_kwargs, eta, seed, progress, only_final, use_graph=True, step_noise=None):
        """x_t, the timestep, the Philox key and all tables stay on the device; one captured graph = `unroll` consecutive
        denoise steps (network evaluation + fused sampler update + timestep decrement), replayed until t reaches 0.

        The loop state (x_t / x0_hat / timestep / seed buffers and the captured graph) is a PLAN kept on the model's
        persistent sampler handle, so every later job of the same shape replays the graph captured by the first one:
        measured on B200, re-capturing per job cost 10 ms typically but 200-1000 ms sporadically (graph instantiate +
        private-pool alloc/free, profiles/r1_job_timeline_before_graph_cache.txt)."""
        device = img.device
        tab = self._tables(device)
        tmap = getattr(self, "timestep_map", list(range(self.num_timesteps)))
        import time as _time
        trace = getattr(self, "trace", None)  # tools/job_timeline.py: list that receives (label, host perf_counter) pairs

        def _tr(label):
            if trace is not None:
                trace.append((label, _time.perf_counter()))
        _tr("loop_begin")
        with th.no_grad():
            from amb200.trace import rng as _nvtx
            with _nvtx("encode_condition"):
                handle = model.sampler_begin(tuple(img.shape), model_kwargs, tmap)
            _tr("sampler_begin_done")
            n = self.num_timesteps
            plans = getattr(handle, "plans", None)
            reuse = plans is not None and use_graph and only_final and step_noise is None
            pkey = (kind, float(eta), int(self.sample_offset), n, id(tab))
            plan = plans.get(pkey) if reuse else None
            if plan is None:
                plan = {"img": th.empty_like(img) if reuse else img, "x0": th.empty_like(img),
                        "t": th.empty(1, device=device, dtype=th.int32), "seed": th.empty(1, device=device, dtype=th.int64),
                        "graph": None, "unroll": 1, "per_graph": 0, "tab": tab}  # `tab` pinned: the graph bakes its pointers
                if reuse:
                    plans[pkey] = plan
            if plan["img"] is not img:
                plan["img"].copy_(img)
                img = plan["img"]
            x0, t_dev, seed_dev = plan["x0"], plan["t"], plan["seed"]
            t_dev.fill_(n - 1)
            seed_dev.fill_(int(seed))

            # CMDM (tc path), ancestral sampling: the sampler update also writes the NEXT step's prologue (bf16 split of x_{t-1}, time
            # token of t-1), so a step is network evaluation + ONE elementwise launch; the first step's prologue runs once per job
            nxt = handle.fuse_next() if (kind == "ddpm" and hasattr(handle, "fuse_next")) else None
            if nxt is not None:
                handle.prepare(img, t_dev)

            def one_step(nz=None):
                with _nvtx("network_eval"):
                    if nxt is not None:
                        handle.forward(img, t_dev, x0, prologue=False)
                    else:
                        handle.forward(img, t_dev, x0)
                if kind == "ddpm":
                    ops.p_sample_update(x0, img, img, nz, tab["coef1"], tab["coef2"], tab["logvar"], t_dev, 0, seed_dev=seed_dev,
                                        sample0=self.sample_offset, nxt=nxt)
                else:
                    ops.ddim_update(x0, img, img, nz, tab["sqrt_recip_ac"], tab["sqrt_recipm1_ac"], tab["ac"], tab["ac_prev"], eta,
                                    t_dev, 0, seed_dev=seed_dev, sample0=self.sample_offset)
                ops.add_i32(t_dev, -1)

            from amb200 import lib as _lib
            launches0, replays, captured_now = _lib.launch_count(), 0, 0
            pbar = None
            if progress:
                from tqdm.auto import tqdm
                pbar = tqdm(total=n)
            k = 0
            while k < n:
                done = 0
                if step_noise is not None:  # parity harness: injected eps, eager
                    one_step(step_noise(k))
                    done = 1
                elif not use_graph:
                    one_step()
                    done = 1
                elif plan["graph"] is None and k == 0:
                    one_step()  # eager first step of the first job: allocates every workspace before capture
                    done = 1
                elif plan["graph"] is None:
                    # one captured graph = `unroll` consecutive denoise steps (fewer host launches per job: the loop is
                    # host-driven, and a stalled host thread would idle the GPU)
                    plan["unroll"] = 8 if (only_final and n - k >= 64) else 1
                    _tr("eager_step_enqueued")
                    th.cuda.synchronize(device)
                    _tr("pre_capture_sync_done")
                    _t0 = _time.perf_counter()
                    graph = th.cuda.CUDAGraph()
                    c0 = _lib.launch_count()
                    with _nvtx("graph_capture"), th.cuda.graph(graph):
                        for _ in range(plan["unroll"]):
                            one_step()
                    plan["per_graph"] = captured_now = _lib.launch_count() - c0
                    plan["graph"] = graph  # capture records without executing: the captured steps still have to run
                    self.last_capture_ms = 1e3 * (_time.perf_counter() - _t0)  # host cost of capture + instantiate
                    _tr("capture_done")
                elif n - k >= plan["unroll"]:
                    with _nvtx("denoise_steps"):
                        plan["graph"].replay()
                    replays += 1
                    done = plan["unroll"]
                else:
                    one_step()  # tail shorter than the captured graph
                    done = 1
                k += done
                if k >= n:
                    _tr("all_enqueued")
                if pbar is not None and done:
                    pbar.update(done)
                self.last_launches = (_lib.launch_count() - launches0 - captured_now) + plan["per_graph"] * replays
                if done and (not only_final or k >= n):
                    last = k >= n and not reuse  # plan buffers are overwritten by the next job: hand out copies
                    yield {"sample": img if last else img.clone(), "pred_xstart": x0 if last else x0.clone()}

    # ------------------------------------------------------------------ training
    def training_losses(self, model, x_start, t, model_kwargs=None, noise=None, **kwargs):
        """:745-826, MSE branch with START_X / EPSILON target; returns {'mse': [B], 'loss': [B]}."""
        self._check_supported()
        if self.loss_type.is_vb():
            raise NotImplementedError("KL losses are outside the reference's configured path (loss_type='MSE')")
        if model_kwargs is None:
            model_kwargs = {}
        x_start = x_start.float().contiguous()
        B = x_start.shape[0]
        x_mask = model_kwargs["x_mask"] if "x_mask" in model_kwargs else th.zeros(x_start.shape[:-1], dtype=th.bool, device=x_start.device)
        if noise is None:
            noise = th.empty_like(x_start)
            ops.randn_(noise, x_start[0].numel(), B, 0, _draw_seed(), 0)
        x_t = self.q_sample(x_start, t, noise=noise)
        model_output = model(x_t, self._scale_timesteps(t), **model_kwargs)
        target = x_start if self.model_mean_type == ModelMeanType.START_X else noise.float()
        assert model_output.shape == target.shape == x_start.shape
        mask3 = x_mask.reshape(B, -1).to(th.uint8).contiguous()
        T = mask3.shape[1]
        if model_output.requires_grad:  # training: differentiable masked MSE (backward kernel am_masked_mse_bwd)
            from amb200.autograd_ops import MaskedMSEFn
            loss = MaskedMSEFn.apply(model_output.float().reshape(B, T, -1), target.reshape(B, T, -1), mask3)
        else:
            loss = th.empty(B, device=x_start.device)
            ops.masked_mse(target.reshape(B, T, -1).contiguous(), model_output.float().reshape(B, T, -1).contiguous(), mask3, loss)
        return {"mse": loss, "loss": loss}


def _extract_into_tensor(arr, timesteps, broadcast_shape):
    """:829-842, kept for API completeness (the hot path never calls it)."""
    res = th.from_numpy(arr).to(device=timesteps.device)[timesteps].float()
    while len(res.shape) < len(broadcast_shape):
        res = res[..., None]
    return res.expand(broadcast_shape)
