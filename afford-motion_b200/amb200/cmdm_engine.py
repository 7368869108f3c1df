"""CUDA execution engine of the CMDM denoiser (models/cmdm.py:118-170,195-196).

Design (B200-first, not a port of the reference's op-by-op graph):
  * conditioning that does not depend on (x_t, t) — text token, 128 contact tokens, key-padding mask — is encoded
    ONCE per batch into a persistent token buffer (the reference recomputes CLIP + the PointTransformer encoder on
    every denoise step, cmdm.py:133-149);
  * the time token is a row gather from a [1000, 512] table built once per weight version;
  * motion_adapter writes straight into the token buffer with the positional encoding fused in its epilogue;
  * every buffer is static per (B, T, G) so one denoise step is CUDA-graph capturable; the timestep lives on device.
"""
import math
import os
from dataclasses import dataclass
from typing import Dict, Optional

import torch

from . import ops
from .pack import c, params_version
from .scene_engine import SceneEncoderEngine


@dataclass
class CMDMCondition:
    """Step-invariant conditioning of one batch."""
    B: int
    G: int
    T: int
    static_tokens: torch.Tensor  # [B, 1+G, D] text + contact tokens, positional encoding already added
    key_pad: Optional[torch.Tensor]  # uint8 [B, S], 1 = ignore key


class CMDMEngine:
    def __init__(self, module):
        self.m = module
        self.scene = SceneEncoderEngine(module.contact_encoder)
        self._version = None
        self.w: Dict[str, torch.Tensor] = {}
        self._ws = {}
        # "tc": tcgen05 3-term bf16-split GEMMs (default); "simt": fp32 SIMT GEMMs (cross-check path)
        self.gemm = os.environ.get("AMB200_GEMM", "tc")
        assert self.gemm in ("tc", "simt")
        # "tc": tcgen05 attention (default with the tc GEMM path, S <= 384); "simt": fp32 SIMT attention kernel
        self.attn = os.environ.get("AMB200_ATTN", "tc")
        # AMB200_LN_FUSE=1: out_proj / linear2 + residual + LayerNorm as one tcgen05 kernel (csrc/gemm_ln_tc.cu).  OFF by default:
        # measured 44.3 / 58.9 us against 34.1 / 42.8 us for GEMM + LayerNorm (profiles/r2_gemm_ln_fused_ab.txt) — its row-per-lane
        # epilogue is bound by uncoalesced residual / output accesses; parity-tested, kept for the next round
        self.fuse_ln = os.environ.get("AMB200_LN_FUSE", "0") == "1"
        # residual add in the LayerNorm's load phase instead of the GEMM epilogue
        self.res_in_ln = os.environ.get("AMB200_RES_IN_LN", "0") == "1"
        # LayerNorms overlapped with the GEMM in front of them through row-block flags (am_linear_tc_set_rowflags / am_layernorm_flags)
        self.ln_overlap = os.environ.get("AMB200_LN_OVERLAP", "0") == "1"
        # last encoder layer on the motion rows only (am_mha_tc_fwd_rows); needs the pipelined attention kernel
        self.last_compact = os.environ.get("AMB200_LAST_COMPACT", "1") == "1" and os.environ.get("AMB200_ATTN_PIPE", "1") != "0"
        assert self.attn in ("tc", "simt")

    # ------------------------------------------------------------------ weights
    def refresh(self):
        v = params_version(self.m)
        if v == self._version:
            return
        m = self.m
        dev = next(m.parameters()).device
        w = {}
        D = m.latent_dim
        pe = m.positional_encoder.pe[:, 0, :].contiguous()  # [5000, D]
        w["pe"] = pe
        # time-token table for every timestep: TimestepEmbedder(t) + PE[0]   (modules.py:52-53, cmdm.py:129,162)
        te = m.timestep_embedder
        pe_t = te.pe[:, 0, :].contiguous()  # [max_len, temb]
        L, temb = pe_t.shape
        h = torch.empty(L, D, device=dev)
        ops.linear(pe_t, c(te.time_embed[0].weight), h, L, D, temb, bias=c(te.time_embed[0].bias), act="silu")
        table = torch.empty(L, D, device=dev)
        ops.linear(h, c(te.time_embed[2].weight), table, L, D, D, bias=c(te.time_embed[2].bias), residual=pe, ldr=D, res_mod=1)
        w["time_table"] = table
        for name in ("language_adapter", "contact_adapter", "motion_adapter", "motion_layer"):
            lin = getattr(m, name)
            w[name + ".w"], w[name + ".b"] = c(lin.weight), c(lin.bias)
        for i, layer in enumerate(m.self_attn_layer.layers):
            p = f"l{i}."
            w[p + "in_w"], w[p + "in_b"] = c(layer.self_attn.in_proj_weight), c(layer.self_attn.in_proj_bias)
            w[p + "out_w"], w[p + "out_b"] = c(layer.self_attn.out_proj.weight), c(layer.self_attn.out_proj.bias)
            w[p + "w1"], w[p + "b1"] = c(layer.linear1.weight), c(layer.linear1.bias)
            w[p + "w2"], w[p + "b2"] = c(layer.linear2.weight), c(layer.linear2.bias)
            w[p + "n1g"], w[p + "n1b"] = c(layer.norm1.weight), c(layer.norm1.bias)
            w[p + "n2g"], w[p + "n2b"] = c(layer.norm2.weight), c(layer.norm2.bias)
            w[p + "eps1"], w[p + "eps2"] = layer.norm1.eps, layer.norm2.eps
        if self.gemm == "tc":  # bf16 (hi|lo) copies of the GEMM weights, K padded to 32
            def sp(t):
                return ops.split_bf16(t, t.shape[0], t.shape[1])
            for name in ("motion_adapter", "motion_layer"):
                w[name + ".w2"] = sp(w[name + ".w"])
            for i in range(len(m.self_attn_layer.layers)):
                for k in ("in_w", "out_w", "w1", "w2"):
                    w[f"l{i}.{k}2"] = sp(w[f"l{i}.{k}"])
        self.w = w
        self.nlayers = len(m.self_attn_layer.layers)
        self.nhead = m.self_attn_layer.layers[0].self_attn.num_heads
        self.ff = m.self_attn_layer.layers[0].linear1.out_features
        self.scene.pack()
        self._version = v

    # ------------------------------------------------------------------ conditioning (once per batch)
    @torch.no_grad()
    def encode_condition(self, text_feat, xyz, contact, x_mask, T, c_text_mask=None, c_text_erase=None, c_pc_mask=None,
                         c_pc_erase=None) -> CMDMCondition:
        self.refresh()
        m, w = self.m, self.w
        dev = xyz.device
        B = xyz.shape[0]
        D = m.latent_dim
        cont = self.scene.forward(xyz, contact)  # [B, G, planes[-1]]
        G, Cc = cont.shape[1], cont.shape[2]
        text = text_feat.float().contiguous()
        if c_text_erase is not None:  # cmdm.py:144-145
            text = (text * (1.0 - c_text_erase.float().view(B, 1))).contiguous()
        if c_pc_erase is not None:  # cmdm.py:154-155
            cont = (cont * (1.0 - c_pc_erase.float().view(B, 1, 1))).contiguous()
        static = torch.empty(B, 1 + G, D, device=dev)
        # text token -> row 0 (+PE[1]); contact tokens -> rows 1..G (+PE[2..1+G])
        ops.linear(text, w["language_adapter.w"], static, B, D, text.shape[1], bias=w["language_adapter.b"],
                   residual=w["pe"][1:2], ldr=D, res_mod=1, ymap=(1, 1 + G, 0))
        ops.linear(cont.view(B * G, Cc), w["contact_adapter.w"], static, B * G, D, Cc, bias=w["contact_adapter.b"],
                   residual=w["pe"][2:2 + G], ldr=D, res_mod=G, ymap=(G, 1 + G, 1))
        key_pad = None
        if m.mask_motion:  # cmdm.py:164-166
            S = 2 + G + T
            kp = torch.zeros(B, S, dtype=torch.bool, device=dev)
            if c_text_mask is not None:
                kp[:, 1] = c_text_mask.view(B).bool()
            if c_pc_mask is not None:
                kp[:, 2:2 + G] = c_pc_mask.view(B, 1).bool()
            kp[:, 2 + G:] = x_mask.bool()
            key_pad = kp.to(torch.uint8).contiguous()
        return CMDMCondition(B=B, G=G, T=T, static_tokens=static, key_pad=key_pad)

    # ------------------------------------------------------------------ workspace
    MAX_WORKSPACES = 4

    def workspace(self, B, G, T, dev):
        """Activation buffers of one (batch, contact tokens, frames) shape.  Keyed on (B, G, T) — not on S = 2+G+T: captured CUDA
        graphs bake these pointers, and two jobs with equal S but different (G, T) must not share (and re-allocate) them.
        Bounded: evicting a workspace also drops the model's sampler handles, whose graphs point into it."""
        key = (B, G, T, str(dev))
        ws = self._ws.get(key)
        if ws is None:
            if len(self._ws) >= self.MAX_WORKSPACES:
                self._ws.clear()
                self.m.__dict__.get("_sampler_handles", {}).clear()
            S = 2 + G + T
            D, ff = self.m.latent_dim, self.ff
            M = B * S
            ws = {
                "X0": torch.empty(B, S, D, device=dev), "Xa": torch.empty(M, D, device=dev), "Xb": torch.empty(M, D, device=dev),
                "QKV": torch.empty(M, 3 * D, device=dev), "ATT": torch.empty(M, D, device=dev), "TMP": torch.empty(M, D, device=dev),
                "Y1": torch.empty(M, D, device=dev), "FF": torch.empty(M, ff, device=dev), "cond_id": None,
            }
            if self.gemm == "tc":
                bf = lambda r, c: torch.zeros(r, c, dtype=torch.bfloat16, device=dev)
                ws.update({"X0S": bf(M, 2 * D), "XSa": bf(M, 2 * D), "XSb": bf(M, 2 * D), "ATTS": bf(M, 2 * D), "Y1S": bf(M, 2 * D),
                           "FFS": bf(M, 2 * ops.pad32(ff)), "xS": bf(B * T, 2 * ops.pad32(self.m.motion_dim)), "QKVS": bf(M, 6 * D),
                           "RS": bf(B * T, 2 * D), "FLAGS": torch.zeros((M + 127) // 128 + 1, dtype=torch.int32, device=dev)})
            self._ws[key] = ws
        return ws

    def bind_condition(self, ws, cond: CMDMCondition):
        """Copy the static tokens into the persistent token buffer (once per batch, not per step)."""
        if ws["cond_id"] is not cond:
            ws["X0"][:, 1:2 + cond.G, :].copy_(cond.static_tokens)
            if self.gemm == "tc":
                D = self.m.latent_dim
                rows = cond.static_tokens.shape[0] * cond.static_tokens.shape[1]
                st2 = ops.split_bf16(cond.static_tokens.view(rows, D), rows, D)
                ws["X0S"].view(cond.B, -1, 2 * D)[:, 1:2 + cond.G, :].copy_(st2.view(cond.B, 1 + cond.G, 2 * D))
            ws["cond_id"] = cond

    # ------------------------------------------------------------------ one network evaluation
    @torch.no_grad()
    def forward(self, x: torch.Tensor, t_dev: torch.Tensor, t_stride: int, cond: CMDMCondition, out: Optional[torch.Tensor] = None,
                time_table: Optional[torch.Tensor] = None, prologue: bool = True):
        """x [B,T,Dm] fp32 contiguous, t_dev int32 device ([1] shared or [B]) -> x0_hat [B,T,Dm].
        prologue=False (device-resident sampling loop, tc path): the time token of this timestep and the bf16 split of x were already
        written by the previous step's fused sampler update (am_p_sample_update_next) or by `step_prologue` at the start of the job."""
        self.refresh()
        w, m = self.w, self.m
        B, T, Dm = x.shape
        G, D = cond.G, m.latent_dim
        S = 2 + G + T
        M = B * S
        ws = self.workspace(B, G, T, x.device)
        self.bind_condition(ws, cond)
        X0 = ws["X0"]
        if self.gemm == "tc":
            return self._forward_tc(x, t_dev, t_stride, cond, out, time_table, ws, prologue)
        ops.gather_time_token(X0, S, D, 0, w["time_table"] if time_table is None else time_table, t_dev, t_stride, B)
        # motion tokens + PE[2+G+j] -> rows 2+G.. of every sample   (cmdm.py:159-162)
        ops.linear(x, w["motion_adapter.w"], X0, B * T, D, Dm, bias=w["motion_adapter.b"], residual=w["pe"][2 + G:2 + G + T], ldr=D,
                   res_mod=T, ymap=(T, S, 2 + G))
        cur = X0.view(M, D)
        H = self.nhead
        hd = D // H
        for i in range(self.nlayers):
            p = f"l{i}."
            nxt = ws["Xa"] if i % 2 == 0 else ws["Xb"]
            ops.linear(cur, w[p + "in_w"], ws["QKV"], M, 3 * D, D, bias=w[p + "in_b"])
            ops.mha_fwd(ws["QKV"], ws["ATT"], cond.key_pad, B, S, H, hd, 1.0 / math.sqrt(hd))
            ops.linear(ws["ATT"], w[p + "out_w"], ws["TMP"], M, D, D, bias=w[p + "out_b"], residual=cur)
            ops.layernorm(ws["TMP"], w[p + "n1g"], w[p + "n1b"], ws["Y1"], M, D, eps=w[p + "eps1"])
            ops.linear(ws["Y1"], w[p + "w1"], ws["FF"], M, self.ff, D, bias=w[p + "b1"], act="gelu")
            ops.linear(ws["FF"], w[p + "w2"], ws["TMP"], M, D, self.ff, bias=w[p + "b2"], residual=ws["Y1"])
            ops.layernorm(ws["TMP"], w[p + "n2g"], w[p + "n2b"], nxt, M, D, eps=w[p + "eps2"])
            cur = nxt
        if out is None:
            out = torch.empty(B, T, Dm, device=x.device)
        ops.linear(cur, w["motion_layer.w"], out, B * T, Dm, D, bias=w["motion_layer.b"], xmap=(T, S, 2 + G))
        return out

    def step_prologue(self, x, t_dev, t_stride, ws, time_table=None):
        """Time token of the current timestep into row 0 of every sample + bf16 (hi|lo) split of x (A operand of the motion adapter)."""
        B, T, Dm = x.shape
        S, D = ws["X0"].shape[1], self.m.latent_dim
        ops.gather_time_token(ws["X0"], S, D, 0, self.w["time_table"] if time_table is None else time_table, t_dev, t_stride, B, x2=ws["X0S"])
        ops.split_bf16(x, B * T, Dm, out=ws["xS"])

    def _forward_tc(self, x, t_dev, t_stride, cond, out, time_table, ws, prologue=True):
        """Same network evaluation with every large GEMM on the tcgen05 path (3-term bf16 split, fp32 accumulate):
        activations travel between GEMMs as bf16 (hi|lo) pairs written by the producing kernel's epilogue
        (GEMM / LayerNorm / attention), fp32 copies are kept only where a residual or the attention kernel needs them."""
        w, m = self.w, self.m
        B, T, Dm = x.shape
        G, D = cond.G, m.latent_dim
        S = 2 + G + T
        M = B * S
        H = self.nhead
        hd = D // H
        X0, X0S = ws["X0"], ws["X0S"]
        Kx = ops.pad32(Dm)
        if prologue:
            self.step_prologue(x, t_dev, t_stride, ws, time_table)
        ops.linear_tc(ws["xS"], w["motion_adapter.w2"], B * T, D, Kx, y=None, y2=X0S, bias=w["motion_adapter.b"],
                      residual=w["pe"][2 + G:2 + G + T], ldr=D, res_mod=T, ymap=(T, S, 2 + G), ldy=D, Np2=D)
        cur, curS = X0.view(M, D), X0S
        ffp = ops.pad32(self.ff)
        fuse_ln = self.fuse_ln and D == 512 and ffp % 64 == 0
        compact = rs_written = False
        ovl = self.ln_overlap and D % 128 == 0 and D % 256 == 0 and not fuse_ln
        will_compact = self.last_compact and self.attn == "tc" and S <= 384 and hd == 64 and not fuse_ln
        for i in range(self.nlayers):
            p = f"l{i}."
            last = i == self.nlayers - 1
            nxt, nxtS = (ws["Xa"], ws["XSa"]) if i % 2 == 0 else (ws["Xb"], ws["XSb"])
            # The last layer's output is only read at the motion tokens (models/cmdm.py:183-186 slices x[non_motion_token:]) and
            # every row is independent after the attention: its attention runs on the query rows [2+G, S) only and writes them
            # compactly, and out_proj / LayerNorm / feed-forward / motion_layer run on B*T rows instead of B*S (60 % at T=196, G=128).
            compact = last and will_compact
            Mr = B * T if compact else M     # rows from the attention output onwards
            resS = curS
            if self.attn == "tc" and S <= 384 and hd == 64:
                ops.linear_tc(curS, w[p + "in_w2"], M, 3 * D, D, y2=ws["QKVS"], bias=w[p + "in_b"], Np2=3 * D)
                ops.mha_tc_fwd(ws["QKVS"], None, ws["ATTS"], cond.key_pad, B, S, H, hd, 1.0 / math.sqrt(hd), q_row0=2 + G if compact else 0)
                if compact:   # the residual stream of the same rows
                    if not rs_written:   # a single-layer trunk: no LayerNorm in front wrote it
                        ws["RS"].view(B, T, 2 * D).copy_(curS.view(B, S, 2 * D)[:, 2 + G:, :])
                    resS = ws["RS"]
            # residual streams travel as the bf16 (hi|lo) pairs the LayerNorm / adapter epilogues already write for the next
            # GEMM's A operand (x = hi + lo, 16 significant bits): no fp32 copy of the activations is written at all
            if fuse_ln:
                # out_proj + residual + norm1 and linear2 + residual + norm2 as ONE tcgen05 kernel each: the full 512-column row sits in
                # the CTA pair's tensor memory, the fp32 hand-off tensor and the LayerNorm launch disappear (csrc/gemm_ln_tc.cu)
                ops.linear_ln_tc(ws["ATTS"], w[p + "out_w2"], M, D, D, w[p + "out_b"], curS, w[p + "n1g"], w[p + "n1b"], w[p + "eps1"], ws["Y1S"])
                ops.linear_tc(ws["Y1S"], w[p + "w12"], M, self.ff, D, y2=ws["FFS"], bias=w[p + "b1"], act="gelu", Np2=ffp)
                ops.linear_ln_tc(ws["FFS"], w[p + "w22"], M, D, ffp, w[p + "b2"], ws["Y1S"], w[p + "n2g"], w[p + "n2b"], w[p + "eps2"], nxtS)
            else:
                win = dict(y2_win=ws["RS"], seg=S, seg_q0=2 + G) if (will_compact and i == self.nlayers - 2) else {}
                if ovl and (Mr + 127) // 128 >= 7:   # (the CTA-pair GEMM needs >= 16 pair tiles)
                    # each LayerNorm starts on the 128-row blocks its GEMM has finished while the GEMM's tail round is still running
                    ops.linear_tc(ws["ATTS"], w[p + "out_w2"], Mr, D, D, y=ws["TMP"], bias=w[p + "out_b"], residual_split=resS, rowflags=ws["FLAGS"])
                    ops.layernorm_flags(ws["TMP"], w[p + "n1g"], w[p + "n1b"], Mr, D, ws["Y1S"], ws["FLAGS"], 4 * D, eps=w[p + "eps1"])
                    ops.linear_tc(ws["Y1S"], w[p + "w12"], Mr, self.ff, D, y2=ws["FFS"], bias=w[p + "b1"], act="gelu", Np2=ffp)
                    ops.linear_tc(ws["FFS"], w[p + "w22"], Mr, D, ffp, y=ws["TMP"], bias=w[p + "b2"], residual_split=ws["Y1S"], rowflags=ws["FLAGS"])
                    ops.layernorm_flags(ws["TMP"], w[p + "n2g"], w[p + "n2b"], Mr, D, nxtS, ws["FLAGS"], 4 * D, eps=w[p + "eps2"], **win)
                    rs_written = rs_written or bool(win)
                    cur, curS = nxt, nxtS
                    continue
                if self.res_in_ln:
                    # the residual add moves from the GEMM epilogue (a TMA load + wait per 16-column half-block, the longest link of
                    # the epilogue chain) into the LayerNorm's load phase: same fp32 sum in the same order, bit-identical
                    ops.linear_tc(ws["ATTS"], w[p + "out_w2"], Mr, D, D, y=ws["TMP"], bias=w[p + "out_b"])
                    ops.layernorm(ws["TMP"], w[p + "n1g"], w[p + "n1b"], None, Mr, D, eps=w[p + "eps1"], y2=ws["Y1S"], residual_split=resS)
                    ops.linear_tc(ws["Y1S"], w[p + "w12"], Mr, self.ff, D, y2=ws["FFS"], bias=w[p + "b1"], act="gelu", Np2=ffp)
                    ops.linear_tc(ws["FFS"], w[p + "w22"], Mr, D, ffp, y=ws["TMP"], bias=w[p + "b2"])
                    ops.layernorm(ws["TMP"], w[p + "n2g"], w[p + "n2b"], None, Mr, D, eps=w[p + "eps2"], y2=nxtS, residual_split=ws["Y1S"], **win)
                    rs_written = rs_written or bool(win)
                    cur, curS = nxt, nxtS
                    continue
                ops.linear_tc(ws["ATTS"], w[p + "out_w2"], Mr, D, D, y=ws["TMP"], bias=w[p + "out_b"], residual_split=resS)
                ops.layernorm(ws["TMP"], w[p + "n1g"], w[p + "n1b"], None, Mr, D, eps=w[p + "eps1"], y2=ws["Y1S"])
                ops.linear_tc(ws["Y1S"], w[p + "w12"], Mr, self.ff, D, y2=ws["FFS"], bias=w[p + "b1"], act="gelu", Np2=ffp)
                ops.linear_tc(ws["FFS"], w[p + "w22"], Mr, D, ffp, y=ws["TMP"], bias=w[p + "b2"], residual_split=ws["Y1S"])
                if will_compact and i == self.nlayers - 2:
                    # the LayerNorm in front of the last layer also writes the motion rows compactly: the last layer's residual stream
                    ops.layernorm(ws["TMP"], w[p + "n2g"], w[p + "n2b"], None, Mr, D, eps=w[p + "eps2"], y2=nxtS, y2_win=ws["RS"], seg=S, seg_q0=2 + G)
                    rs_written = True
                else:
                    ops.layernorm(ws["TMP"], w[p + "n2g"], w[p + "n2b"], None, Mr, D, eps=w[p + "eps2"], y2=nxtS)
            cur, curS = nxt, nxtS
        if out is None:
            out = torch.empty(B, T, Dm, device=x.device)
        if self.nlayers and compact:   # curS already holds the B*T motion rows
            ops.linear_tc(curS, w["motion_layer.w2"], B * T, Dm, D, y=out, bias=w["motion_layer.b"], ldy=Dm)
        else:   # motion_layer over all tokens; only rows s >= 2+G are written (skip map) -> out [B*T, Dm]
            ops.linear_tc(curS, w["motion_layer.w2"], M, Dm, D, y=out, bias=w["motion_layer.b"], ymap=(S, T, -(2 + G)), ldy=Dm)
        return out
