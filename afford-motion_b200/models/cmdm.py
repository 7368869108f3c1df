"""CMDM (paper: AMDM) — drop-in for /root/reference/models/cmdm.py (arch='trans_enc').

Same class name / registry entry, constructor `CMDM(cfg.model, device=...)`, `forward(x, timesteps, **kwargs)`,
config keys (configs/model/cmdm.yaml) and state_dict names; the arithmetic runs in amb200.cmdm_engine (CUDA).
`arch='trans_dec'` is out of scope (no reference script uses it).
"""
import torch
import torch.nn as nn

from models.base import Model
from models.functions import encode_text_clip, get_lang_feat_dim_type, load_and_freeze_clip_model
from models.modules import PositionalEncoding, SceneMapEncoder, TimestepEmbedder

_REPR_DIM = {"smplx_no_hands": 69, "pos": 66, "pos_rot": 129, "contact_one_joints": 1, "contact_all_joints": 22,
             "contact_cont_joints": 6, "contact_pelvis": 1, "h3d": 263}


def compute_repr_dimesion(data_repr: str) -> int:
    """utils/misc.py:4-22 (kept local: importing the reference's utils.misc instantiates an SMPL-X layer)."""
    if data_repr not in _REPR_DIM:
        raise ValueError(f"Unknown data representation: {data_repr}")
    return _REPR_DIM[data_repr]


@Model.register()
class CMDM(nn.Module):
    def __init__(self, cfg, *args, **kwargs):
        super().__init__()
        self.device = kwargs["device"] if "device" in kwargs else "cpu"
        self.motion_type = cfg.data_repr
        self.motion_dim = cfg.input_feats
        self.latent_dim = cfg.latent_dim
        self.mask_motion = cfg.mask_motion
        self.arch = cfg.arch
        if self.arch != "trans_enc":
            raise NotImplementedError("afford-motion_b200 implements CMDM arch='trans_enc' (the only one the reference scripts use)")

        self.time_emb_dim = cfg.time_emb_dim
        self.timestep_embedder = TimestepEmbedder(self.latent_dim, self.time_emb_dim, max_len=1000)

        self.contact_type = cfg.contact_model.contact_type
        self.contact_dim = compute_repr_dimesion(self.contact_type)
        self.planes = list(cfg.contact_model.planes)
        self.contact_adapter = nn.Linear(self.planes[-1], self.latent_dim, bias=True)
        self.contact_encoder = SceneMapEncoder(point_feat_dim=self.contact_dim, planes=self.planes,
                                               blocks=list(cfg.contact_model.blocks), num_points=cfg.contact_model.num_points)

        self.text_model_name = cfg.text_model.version
        self.text_max_length = cfg.text_model.max_length
        self.text_feat_dim, self.text_feat_type = get_lang_feat_dim_type(self.text_model_name)
        if self.text_feat_type != "clip":
            raise NotImplementedError("only CLIP text features are supported (every reference script uses ViT-B/32)")
        self.text_model = load_and_freeze_clip_model(self.text_model_name)
        self.language_adapter = nn.Linear(self.text_feat_dim, self.latent_dim, bias=True)

        self.motion_adapter = nn.Linear(self.motion_dim, self.latent_dim, bias=True)
        self.positional_encoder = PositionalEncoding(self.latent_dim, dropout=0.1, max_len=5000)
        self.num_layers = list(cfg.num_layers)
        # parameter container with torch's packed in_proj layout (checkpoint compatibility, cmdm.py:66-77)
        self.self_attn_layer = nn.TransformerEncoder(
            nn.TransformerEncoderLayer(d_model=self.latent_dim, nhead=cfg.num_heads, dim_feedforward=cfg.dim_feedforward,
                                       dropout=cfg.dropout, activation="gelu", batch_first=True),
            enable_nested_tensor=False, num_layers=sum(self.num_layers))
        self.motion_layer = nn.Linear(self.latent_dim, self.motion_dim, bias=True)

        self._engine = None
        self._cond_cache = None

    # ------------------------------------------------------------------ engine plumbing
    @property
    def engine(self):
        if self._engine is None:
            from amb200.cmdm_engine import CMDMEngine
            self._engine = CMDMEngine(self)
        return self._engine

    _COND_TENSOR_KEYS = ("c_pc_xyz", "c_pc_contact", "x_mask", "c_text_mask", "c_text_erase", "c_pc_mask", "c_pc_erase")

    def _cond_lookup(self, kwargs):
        """Per-step `forward()` cache (the reference's loop calls the model with the SAME kwargs dict on every step, test.py:94-101).
        Keyed on the identity of the input tensor OBJECTS, which the cache entry keeps alive (so an address can never be recycled
        under it), plus their in-place-modification counters — never on data_ptr."""
        ents = tuple(kwargs.get(k) for k in self._COND_TENSOR_KEYS)
        sig = (tuple(kwargs["c_text"]), tuple(None if t is None else t._version for t in ents), self.engine._version)
        c = self._cond_cache
        if c is not None and c["sig"] == sig and all(a is b for a, b in zip(c["refs"], ents)):
            return c["cond"], ents, sig
        return None, ents, sig

    def encode_condition(self, T, use_cache=True, **kwargs):
        """Step-invariant conditioning (text token, contact tokens, key mask).  use_cache=False (sampler_begin: once per job)
        always re-encodes."""
        self.engine.refresh()
        ents = sig = None
        if use_cache:
            cond, ents, sig = self._cond_lookup(kwargs)
            if cond is not None:
                return cond
        dev = kwargs["c_pc_xyz"].device
        text = encode_text_clip(self.text_model, kwargs["c_text"], max_length=self.text_max_length, device=dev).detach().float()
        B = kwargs["c_pc_xyz"].shape[0]
        x_mask = kwargs.get("x_mask")
        if x_mask is None:
            x_mask = torch.zeros(B, T, dtype=torch.bool, device=dev)
        cond = self.engine.encode_condition(text, kwargs["c_pc_xyz"], kwargs["c_pc_contact"], x_mask, T,
                                            c_text_mask=kwargs.get("c_text_mask"), c_text_erase=kwargs.get("c_text_erase"),
                                            c_pc_mask=kwargs.get("c_pc_mask"), c_pc_erase=kwargs.get("c_pc_erase"))
        if use_cache:
            self._cond_cache = dict(sig=sig, refs=ents, cond=cond)
        return cond

    def forward(self, x, timesteps, **kwargs):
        """x [bs, seq_len, motion_dim], timesteps int64 [bs] -> [bs, seq_len, motion_dim]  (cmdm.py:118-196)."""
        if not x.is_cuda:
            raise RuntimeError("afford-motion_b200: CMDM runs on CUDA (sm_100a) only — there is no CPU fallback")
        if self.training:
            # training path (utils/training.py:141-154): autograd graph of libamb200 kernels, batch-statistics BatchNorm,
            # dropout, conditioning re-encoded every step
            from amb200.cmdm_train import cmdm_forward_train
            text = encode_text_clip(self.text_model, kwargs["c_text"], max_length=self.text_max_length, device=x.device).detach().float()
            return cmdm_forward_train(self, x.float().contiguous(), timesteps, text, kwargs)
        cond = self.encode_condition(x.shape[1], **kwargs)
        _check_timesteps(timesteps, self.engine.w["time_table"].shape[0])
        t_dev = timesteps.to(device=x.device, dtype=torch.int32).contiguous()
        return self.engine.forward(x.float().contiguous(), t_dev, 1, cond)

    def sampler_begin(self, shape, model_kwargs, timestep_map):
        """Device-resident sampling hook used by diffusion.gaussian_diffusion._fast_loop: conditioning encoded once per job and
        bound into the engine's persistent token buffer; the handle itself (re-indexed time-token table, key-padding buffer,
        loop plans with their captured CUDA graph) is PERSISTENT per (shape, timestep map, weight version), so later jobs of
        the same shape replay the first job's graph."""
        cond = self.encode_condition(shape[1], use_cache=False, **model_kwargs)  # once per job, never from a cache
        eng = self.engine
        from amb200 import lib as _lib
        key = (tuple(shape), tuple(timestep_map), eng._version, _lib.get_precision(), cond.key_pad is None, cond.G, str(cond.static_tokens.device))
        handles = self.__dict__.setdefault("_sampler_handles", {})
        h = handles.get(key)
        if h is None:
            if len(handles) >= 8:  # each handle pins a captured graph and its memory pool
                handles.clear()
            h = handles[key] = _CMDMSamplerHandle(eng, cond, timestep_map)
        h.rebind(cond)
        return h


def _check_timesteps(timesteps, table_rows):
    """models/modules.py:50 indexes pe[timesteps]: out-of-range steps raise there; the kernels index device tables unchecked, so
    the per-call `forward()` path validates on the host (one small D2H read; the device-resident sampling loop validates its
    timestep map once per handle instead)."""
    if timesteps.numel() and (int(timesteps.max()) >= table_rows or int(timesteps.min()) < 0):
        raise IndexError(f"timestep out of range for the TimestepEmbedder table ({table_rows} rows)")


class _CMDMSamplerHandle:
    """Persistent per-(shape, timestep map, weight version) sampling state; see CMDM.sampler_begin."""

    def __init__(self, eng, cond, timestep_map):
        from amb200.cmdm_engine import CMDMCondition
        dev = cond.static_tokens.device
        if max(timestep_map) >= eng.w["time_table"].shape[0] or min(timestep_map) < 0:  # the reference's pe[t] raises the same way
            raise IndexError(f"timestep {max(timestep_map)} is out of range for the TimestepEmbedder table ({eng.w['time_table'].shape[0]} rows)")
        idx = torch.as_tensor(list(timestep_map), dtype=torch.long).to(dev)  # once per handle (a pageable H2D copy synchronises)
        self.eng = eng
        self.table = eng.w["time_table"][idx].contiguous()
        self.cond = CMDMCondition(B=cond.B, G=cond.G, T=cond.T, static_tokens=cond.static_tokens,
                                  key_pad=None if cond.key_pad is None else torch.empty_like(cond.key_pad))
        self.plans = {}

    def rebind(self, cond):
        """New job: copy its conditioning into the buffers the captured graph reads."""
        ws = self.eng.workspace(cond.B, cond.G, cond.T, cond.static_tokens.device)
        ws["cond_id"] = None
        self.eng.bind_condition(ws, cond)
        if self.cond.key_pad is not None:
            self.cond.key_pad.copy_(cond.key_pad)
        self.cond.static_tokens = cond.static_tokens
        ws["cond_id"] = self.cond

    def forward(self, x, t_dev, out, prologue=True):
        return self.eng.forward(x, t_dev, 0, self.cond, out=out, time_table=self.table, prologue=prologue)

    def _ws(self):
        c = self.cond
        return self.eng.workspace(c.B, c.G, c.T, c.static_tokens.device)

    def fuse_next(self):
        """Buffers the fused sampler update writes for the next step (ops.p_sample_update(nxt=...)); None on the SIMT GEMM path."""
        if self.eng.gemm != "tc":
            return None
        from amb200 import ops
        ws, m = self._ws(), self.eng.m
        return dict(xs2=ws["xS"], D=m.motion_dim, Kx=ops.pad32(m.motion_dim), tokX=ws["X0"], tokX2=ws["X0S"], S=ws["X0"].shape[1],
                    TD=m.latent_dim, table=self.table)

    def prepare(self, x, t_dev):
        """Once per job, before the first step: what every later step gets from the previous step's fused update."""
        self.eng.step_prologue(x, t_dev, 0, self._ws(), self.table)
