"""CUDA execution engine of the CDM ContactPerceiver denoiser (models/cdm.py:155-188,474-513).

The Perceiver here has exactly 2 latent tokens (text, time), so both cross-attentions are folded exactly
 instead of materialising K/V [B,N,512] like the reference:
  encoder : score = (W_k^T q).LN(e) + q.b_k,  sum_j p_j V_j = W_v (sum_j p_j LN(e_j)) + b_v     (am_cdm_encoder_*)
  decoder : decoder_adapter∘encoder_adapter is ONE 9->256 affine map; q_proj / o_proj fold against the 2 K/V tokens
            (am_cdm_decoder_point); the second MLP linear folds into the 256->6 output head (am_linear_skinny).
What remains dense per point is the 256->256 GELU layer (one GEMM over B*N rows).  The latent side (2 tokens per
sample) is a chain of small GEMM / LayerNorm launches, graph-captured with the rest of the step.
"""
import os
from dataclasses import dataclass
from typing import Dict, Optional

import torch

from . import ops
from .cdm_fold import fold_constants
from .pack import c, params_version

R = 16  # (head, latent) rows


@dataclass
class CDMCondition:
    B: int
    N: int
    xyz: torch.Tensor          # [B,N,3]
    text_latent: torch.Tensor  # [B,512] language_adapter(text)
    point_feat: Optional[torch.Tensor]


def _d(t):
    return t.detach().double()


class CDMEngine:
    def __init__(self, module):
        self.m = module
        self._version = None
        self.w: Dict[str, torch.Tensor] = {}
        self._ws = {}
        self.gemm = os.environ.get("AMB200_GEMM", "tc")  # tcgen05 (default) or fp32 SIMT for the per-point MLP GEMM
        # point path: "collapsed" (default when cin == 9: rank-collapsed kernels of csrc/perceiver_tc.cu) or "general"
        self.point_path = os.environ.get("AMB200_CDM_POINTS", "collapsed")
        # latent side of the collapsed path: "fused" (two cluster kernels, csrc/perceiver_latent.cu) or "layers" (one launch per layer)
        self.latent_path = os.environ.get("AMB200_CDM_LATENT", "fused")
        self.K = None
        self.latw = None

    def refresh(self):
        v = params_version(self.m)
        if v == self._version:
            return
        m = self.m
        cm = m.contact_model
        dev = next(m.parameters()).device
        w = {}
        # latent time table: time_embedding_adapter(TimestepEmbedder(t)) for every t   (cdm.py:176-177,485)
        te = m.timestep_embedder
        pe_t = te.pe[:, 0, :].contiguous()
        L, temb = pe_t.shape
        dm = te.d_model
        h = torch.empty(L, dm, device=dev)
        ops.linear(pe_t, c(te.time_embed[0].weight), h, L, dm, temb, bias=c(te.time_embed[0].bias), act="silu")
        h2 = torch.empty(L, dm, device=dev)
        ops.linear(h, c(te.time_embed[2].weight), h2, L, dm, dm, bias=c(te.time_embed[2].bias))
        DL = cm.time_embedding_adapter.out_features
        table = torch.empty(L, DL, device=dev)
        ops.linear(h2, c(cm.time_embedding_adapter.weight), table, L, DL, dm, bias=c(cm.time_embedding_adapter.bias))
        w["time_table"] = table
        w["la_w"], w["la_b"] = c(cm.language_adapter.weight), c(cm.language_adapter.bias)
        self.DL = DL

        # ---- encoder cross-attention (modules.py:504-541)
        ca = cm.encoder_cross_attn[0].module
        att = ca.attention
        H = att.num_heads
        self.He = H
        hd = DL // H
        scale = hd ** -0.5
        w["e_qn_g"], w["e_qn_b"] = c(ca.q_norm.weight), c(ca.q_norm.bias)
        w["e_kvn_g"], w["e_kvn_b"] = c(ca.kv_norm.weight), c(ca.kv_norm.bias)
        w["e_q_w"], w["e_q_b"] = c(att.q_proj.weight * scale), c(att.q_proj.bias * scale)  # q *= dp_scale (:335)
        C = att.k_proj.in_features
        self.C = C
        # folded key matrices: per head [C+1, hd]: rows 0..C-1 = Wk_h^T, row C = bk_h
        wk = att.k_proj.weight.detach().view(H, hd, C)
        bk = att.k_proj.bias.detach().view(H, hd)
        w["e_kfold"] = c(torch.cat([wk.transpose(1, 2), bk.unsqueeze(1)], dim=1))  # [H, C+1, hd]
        w["e_v_w"], w["e_v_b"] = c(att.v_proj.weight), c(att.v_proj.bias)  # rows h*hd.. = head h
        w["e_o_w"], w["e_o_b"] = c(att.o_proj.weight), c(att.o_proj.bias)
        mlp = cm.encoder_cross_attn[1].module
        w["e_m_g"], w["e_m_b"] = c(mlp[0].weight), c(mlp[0].bias)
        w["e_m1_w"], w["e_m1_b"] = c(mlp[1].weight), c(mlp[1].bias)
        w["e_m2_w"], w["e_m2_b"] = c(mlp[3].weight), c(mlp[3].bias)
        w["ea_w"], w["ea_b"] = c(cm.encoder_adapter.weight), c(cm.encoder_adapter.bias)

        # ---- latent self-attention layers (modules.py:544-648)
        self.n_self = len(cm.encoder_self_attn)
        for i, layer in enumerate(cm.encoder_self_attn):
            sa = layer[0].module
            a = sa.attention
            p = f"s{i}_"
            w[p + "n_g"], w[p + "n_b"] = c(sa.norm.weight), c(sa.norm.bias)
            w[p + "qkv_w"] = c(torch.cat([a.q_proj.weight, a.k_proj.weight, a.v_proj.weight], 0))
            w[p + "qkv_b"] = c(torch.cat([a.q_proj.bias, a.k_proj.bias, a.v_proj.bias], 0))
            w[p + "o_w"], w[p + "o_b"] = c(a.o_proj.weight), c(a.o_proj.bias)
            ml = layer[1].module
            w[p + "m_g"], w[p + "m_b"] = c(ml[0].weight), c(ml[0].bias)
            w[p + "m1_w"], w[p + "m1_b"] = c(ml[1].weight), c(ml[1].bias)
            w[p + "m2_w"], w[p + "m2_b"] = c(ml[3].weight), c(ml[3].bias)

        # ---- decoder cross-attention, folded against the 2 latent K/V tokens
        dc = cm.decoder_cross_attn[0].module
        da = dc.attention
        Hd = da.num_heads
        self.Hd = Hd
        Cq = da.q_proj.in_features
        assert Cq == C == 256, "CUDA kernels are specialised for 256 point channels (configs/model/cdm.yaml:38-46)"
        hdd = da.q_proj.out_features // Hd
        dscale = hdd ** -0.5
        w["d_qn_g"], w["d_qn_b"] = c(dc.q_norm.weight), c(dc.q_norm.bias)
        w["d_kvn_g"], w["d_kvn_b"] = c(dc.kv_norm.weight), c(dc.kv_norm.bias)
        w["d_kv_w"] = c(torch.cat([da.k_proj.weight, da.v_proj.weight], 0))  # [2*Cq, DL]
        w["d_kv_b"] = c(torch.cat([da.k_proj.bias, da.v_proj.bias], 0))
        wq = da.q_proj.weight.detach().view(Hd, hdd, Cq) * dscale
        bq = da.q_proj.bias.detach().view(Hd, hdd) * dscale
        w["d_qfold"] = c(torch.cat([wq.transpose(1, 2), bq.unsqueeze(1)], dim=1))  # [Hd, Cq+1, hdd]
        w["d_o_w"], w["d_o_b"] = c(da.o_proj.weight), c(da.o_proj.bias)
        self.hdd = hdd
        # decoder_adapter ∘ encoder_adapter : one (cin -> 256) affine map, folded in fp64
        Wd = _d(cm.decoder_adapter.weight) @ _d(cm.encoder_adapter.weight)
        bd = _d(cm.decoder_adapter.weight) @ _d(cm.encoder_adapter.bias) + _d(cm.decoder_adapter.bias)
        w["dd_w"], w["dd_b"] = Wd.float().contiguous(), bd.float().contiguous()
        dm_ = cm.decoder_cross_attn[1].module
        w["d_m_g"], w["d_m_b"] = c(dm_[0].weight), c(dm_[0].bias)
        w["d_m1_w"], w["d_m1_b"] = c(dm_[1].weight), c(dm_[1].bias)
        if self.gemm == "tc":
            w["d_m1_w2"] = ops.split_bf16(w["d_m1_w"], w["d_m1_w"].shape[0], w["d_m1_w"].shape[1])
        # contact_layer(h1 + W2 g + b2) = [Wc | Wc W2] [h1; g] + (Wc b2 + bc)
        Wc, bc = _d(m.contact_layer.weight), _d(m.contact_layer.bias)
        w["head_w"] = torch.cat([Wc, Wc @ _d(dm_[3].weight)], dim=1).float().contiguous()
        w["head_b"] = (Wc @ _d(dm_[3].bias) + bc).float().contiguous()
        self.out_dim = m.contact_layer.out_features
        # rank-collapsed point path (H3D configs: u = cat(x_t [6], xyz [3]); see amb200/cdm_fold.py)
        self.K = None
        if cm.encoder_adapter.in_features == 9 and self.out_dim == 6 and self.He == 8 and self.Hd == 8:
            self.K = {k: (t.to(dev) if torch.is_tensor(t) else t) for k, t in fold_constants(m).items()}
        self.w = w
        self.latw = None
        if self.K is not None and DL == 512 and C == 256 and self.n_self == 2 and hd == 64 and hdd == 32 and mlp[1].out_features == DL:
            self.latw = self._latent_table(w, self.K)
        self._version = v

    def _latent_table(self, w, K):
        """HOST pointer table of the fused latent-chain kernels (order of `struct LatW`, csrc/perceiver_latent.cu); weights are
        re-laid out K-major ([K][N]) once per weight version.  Returns (ctypes array, count, keep-alive tensors)."""
        import ctypes
        from . import lib as _l
        T = lambda t: t.t().contiguous()  # noqa: E731
        order = [w["la_w"],  # unused slot (keeps the table aligned with the struct)
                 w["e_qn_g"], w["e_qn_b"], T(w["e_q_w"]), w["e_q_b"], K["e_kfold"], K["e_ecg"], K["e_beta"],
                 T(w["e_v_w"]), w["e_v_b"], T(w["e_o_w"]), w["e_o_b"],
                 w["e_m_g"], w["e_m_b"], T(w["e_m1_w"]), w["e_m1_b"], T(w["e_m2_w"]), w["e_m2_b"]]
        for name in ("n_g", "n_b", "qkv_w", "qkv_b", "o_w", "o_b", "m_g", "m_b", "m1_w", "m1_b", "m2_w", "m2_b"):
            for i in range(2):
                t = w[f"s{i}_{name}"]
                order.append(T(t) if name.endswith("_w") else t)
        order += [w["d_kvn_g"], w["d_kvn_b"], T(w["d_kv_w"]), w["d_kv_b"], K["d_qfold"], T(K["d_ostack"])]
        n = int(_l.load().am_cdm_latent_nweights())
        assert len(order) == n, (len(order), n)
        assert all(t.dtype == torch.float32 and t.is_contiguous() and t.is_cuda for t in order)
        arr = (ctypes.c_void_p * n)(*[t.data_ptr() for t in order])
        return arr, n, order

    # ------------------------------------------------------------------ conditioning (once per batch)
    @torch.no_grad()
    def encode_condition(self, text_feat, xyz, point_feat=None) -> CDMCondition:
        self.refresh()
        B, N, _ = xyz.shape
        text = text_feat.float().contiguous()
        lat = torch.empty(B, self.DL, device=xyz.device)
        ops.linear(text, self.w["la_w"], lat, B, self.DL, text.shape[1], bias=self.w["la_b"])
        pf = None if point_feat is None else point_feat.float().contiguous()
        return CDMCondition(B=B, N=N, xyz=xyz.float().contiguous(), text_latent=lat, point_feat=pf)

    def workspace(self, B, N, dev):
        key = (B, N, str(dev))
        ws = self._ws.get(key)
        if ws is None:
            self._evict()
            DL, C = self.DL, self.C
            e = lambda *s: torch.empty(*s, device=dev)
            nchunk = 1
            while B * nchunk < 2 * 148 and N // (nchunk * 2) >= 64:
                nchunk *= 2
            ws = dict(L0=e(B, 2, DL), LN=e(2 * B, DL), Q=e(2 * B, DL), QF=torch.zeros(B, R, C + 4, device=dev), Z=e(B, R, C),
                      PART=e(B, nchunk, R, C + 2), AO=e(2 * B, DL), La=e(2 * B, DL), Lb=e(2 * B, DL), Hh=e(2 * B, DL),
                      QKV=e(2 * B, 3 * DL), KV=e(2 * B, 2 * C), KF=torch.zeros(B, R, C + 4, device=dev), U=e(B, R, C), H1=e(B * N, C),
                      HN=None if self.gemm == "tc" else e(B * N, C), HN2=torch.zeros(B * N, 2 * C, dtype=torch.bfloat16, device=dev) if self.gemm == "tc" else None,
                      Gm=e(B * N, C), nchunk=nchunk, cond_id=None)
            self._ws[key] = ws
        return ws

    def _evict(self, limit=4):
        """Bound the workspace cache (the general path holds ~1.5 GB per (64, 8192) shape).  Captured graphs bake workspace
        pointers, so evicting also drops the model's sampler handles."""
        if len(self._ws) >= limit:
            self._ws.clear()
            self.m.__dict__.get("_sampler_handles", {}).clear()

    def workspace_collapsed(self, B, N, dev):
        key = ("c", B, N, str(dev))
        ws = self._ws.get(key)
        if ws is None:
            self._evict()
            DL, C = self.DL, self.C
            d = self.K["dims"]
            AEW, NS = d["KU"] + 2, d["NS"]
            e = lambda *s: torch.empty(*s, device=dev)
            nchunk = 1
            while B * nchunk < 2 * 148 and N // (nchunk * 2) >= 256:
                nchunk *= 2
            ws = dict(L0=e(B, 2, DL), LN=e(2 * B, DL), Q=e(2 * B, DL), AE=e(B, R, AEW), PART=e(B, nchunk, R, AEW), Z=e(B, R, C),
                      AO=e(2 * B, DL), La=e(2 * B, DL), Lb=e(2 * B, DL), Hh=e(2 * B, DL), QKV=e(2 * B, 3 * DL), KV=e(2 * B, 2 * C),
                      AQ=e(B, R, AEW), UU=e(B, R, NS), PB=torch.zeros(B, 768, device=dev),
                      BLOB=torch.zeros(B, 32768, dtype=torch.uint8, device=dev), nchunk=nchunk, cond_id=None)
            self._ws[key] = ws
        return ws

    def is_collapsed(self, cond) -> bool:
        return self.K is not None and cond.point_feat is None and self.point_path == "collapsed"

    def workspace_for(self, cond):
        dev = cond.xyz.device
        return self.workspace_collapsed(cond.B, cond.N, dev) if self.is_collapsed(cond) else self.workspace(cond.B, cond.N, dev)

    @torch.no_grad()
    def forward(self, x, t_dev, t_stride, cond: CDMCondition, out=None, time_table=None):
        """x [B,N,cx] fp32, t_dev int32 device -> x0_hat [B,N,contact_dim]."""
        self.refresh()
        w = self.w
        B, N, cx = x.shape
        dev = x.device
        DL, C, He, Hd = self.DL, self.C, self.He, self.Hd
        if cond.point_feat is not None:  # cdm.py:167-168
            x = torch.cat([x, cond.point_feat], dim=-1).contiguous()
            cx = x.shape[-1]
        collapsed = self.is_collapsed(cond)
        ws = self.workspace_for(cond)
        if collapsed and self.latw is not None and self.latent_path == "fused":
            K = self.K
            NS = K["dims"]["NS"]
            tt = w["time_table"] if time_table is None else time_table
            if out is None:
                out = torch.empty(B, N, self.out_dim, device=dev)
            ops.cdm_latent_pre(self.latw, cond.text_latent, tt, t_dev, t_stride, ws["AE"], B)
            ops.cdm_enc_points(x, cond.xyz, K["e_chol"], ws["AE"], ws["PART"], B, N, ws["nchunk"])
            ops.cdm_latent_post(self.latw, cond.text_latent, tt, t_dev, t_stride, ws["PART"], ws["nchunk"], ws["AQ"], ws["UU"], NS, B)
            ops.cdm_dec_prep(ws["AQ"], ws["UU"], NS, K["d_g1uu"], K["d_mu"], K["d_hu"], ws["PB"], ws["BLOB"], B)
            ops.cdm_dec_points_tc(x, cond.xyz, K["d_chol"], K["d_c1"], K["d_wg"], ws["PB"], ws["BLOB"], out, B, N)
            return out
        if ws["cond_id"] is not cond:
            ws["L0"][:, 0, :].copy_(cond.text_latent)
            ws["cond_id"] = cond
        L0 = ws["L0"]
        ops.gather_time_token(L0, 2, DL, 1, w["time_table"] if time_table is None else time_table, t_dev, t_stride, B)
        M2 = 2 * B
        # ---- encoder cross-attention
        ops.layernorm(L0, w["e_qn_g"], w["e_qn_b"], ws["LN"], M2, DL)
        ops.linear(ws["LN"], w["e_q_w"], ws["Q"], M2, DL, DL, bias=w["e_q_b"])
        hd = DL // He
        if collapsed:
            K = self.K
            AEW = K["dims"]["KU"] + 2
            # AE[b, 2h+l, :] = q_{h,l}^T [Wk_h diag(g) Ec | Wk_h beta + bk_h]      (all heads in one batched launch)
            ops.linear_batched(ws["Q"], K["e_kfold"], ws["AE"], M2, AEW, hd, He, hd, AEW * hd, 2 * AEW, ldx=DL, ldy=AEW, ymap=(2, R, 0))
            ops.cdm_enc_points(x, cond.xyz, K["e_chol"], ws["AE"], ws["PART"], B, N, ws["nchunk"])
            ops.cdm_enc_expand(ws["PART"], K["e_ecg"], K["e_beta"], ws["Z"], B, ws["nchunk"])
        else:
            # qf[b, 2h+l, :C] = Wk_h^T q_{h,l} ; [.., C] = q_{h,l}.bk_h      (all heads in one batched launch)
            ops.linear_batched(ws["Q"], w["e_kfold"], ws["QF"], M2, C + 1, hd, He, hd, (C + 1) * hd, 2 * (C + 4), ldx=DL, ldy=C + 4,
                               ymap=(2, R, 0))
            ops.cdm_encoder_partial(x, cond.xyz, w["ea_w"], w["ea_b"], w["e_kvn_g"], w["e_kvn_b"], ws["QF"], C + 4, ws["PART"], B, N, cx,
                                    ws["nchunk"])
            ops.cdm_encoder_combine(ws["PART"], ws["Z"], B, ws["nchunk"])
        # o_{h,l} = Wv_h z_{h,l} + bv_h
        ops.linear_batched(ws["Z"], w["e_v_w"], ws["AO"], M2, hd, C, He, 2 * C, hd * C, hd, bias=w["e_v_b"], bb=hd, ldy=DL, xmap=(2, R, 0))
        ops.linear(ws["AO"], w["e_o_w"], ws["La"], M2, DL, DL, bias=w["e_o_b"], residual=L0)  # residual on un-normalised L (:230)
        ops.layernorm(ws["La"], w["e_m_g"], w["e_m_b"], ws["LN"], M2, DL)
        ops.linear(ws["LN"], w["e_m1_w"], ws["Hh"], M2, DL, DL, bias=w["e_m1_b"], act="gelu")
        ops.linear(ws["Hh"], w["e_m2_w"], ws["Lb"], M2, DL, DL, bias=w["e_m2_b"], residual=ws["La"])
        cur, oth = ws["Lb"], ws["La"]
        # ---- latent self-attention
        for i in range(self.n_self):
            p = f"s{i}_"
            ops.layernorm(cur, w[p + "n_g"], w[p + "n_b"], ws["LN"], M2, DL)
            ops.linear(ws["LN"], w[p + "qkv_w"], ws["QKV"], M2, 3 * DL, DL, bias=w[p + "qkv_b"])
            ops.mha_fwd(ws["QKV"], ws["AO"], None, B, 2, He, hd, hd ** -0.5)
            ops.linear(ws["AO"], w[p + "o_w"], oth, M2, DL, DL, bias=w[p + "o_b"], residual=cur)
            ops.layernorm(oth, w[p + "m_g"], w[p + "m_b"], ws["LN"], M2, DL)
            ops.linear(ws["LN"], w[p + "m1_w"], ws["Hh"], M2, DL, DL, bias=w[p + "m1_b"], act="gelu")
            ops.linear(ws["Hh"], w[p + "m2_w"], cur, M2, DL, DL, bias=w[p + "m2_b"], residual=oth)
        # ---- decoder: fold q_proj / o_proj against the 2 latent K/V tokens
        ops.layernorm(cur, w["d_kvn_g"], w["d_kvn_b"], ws["LN"], M2, DL)
        ops.linear(ws["LN"], w["d_kv_w"], ws["KV"], M2, 2 * C, DL, bias=w["d_kv_b"])
        hdd = self.hdd
        if out is None:
            out = torch.empty(B, N, self.out_dim, device=dev)
        if collapsed:
            NS = K["dims"]["NS"]
            ops.linear_batched(ws["KV"], K["d_qfold"], ws["AQ"], M2, AEW, hdd, Hd, hdd, AEW * hdd, 2 * AEW, ldx=2 * C, ldy=AEW,
                               ymap=(2, R, 0))
            ops.linear_batched(ws["KV"][:, C:], K["d_ostack"], ws["UU"], M2, NS, hdd, Hd, hdd, hdd, 2 * NS, ldx=2 * C, ldw=C, ldy=NS,
                               ymap=(2, R, 0))
            ops.cdm_dec_prep(ws["AQ"], ws["UU"], NS, K["d_g1uu"], K["d_mu"], K["d_hu"], ws["PB"], ws["BLOB"], B)
            ops.cdm_dec_points_tc(x, cond.xyz, K["d_chol"], K["d_c1"], K["d_wg"], ws["PB"], ws["BLOB"], out, B, N)
            return out
        ops.linear_batched(ws["KV"], w["d_qfold"], ws["KF"], M2, C + 1, hdd, Hd, hdd, (C + 1) * hdd, 2 * (C + 4), ldx=2 * C, ldy=C + 4,
                           ymap=(2, R, 0))
        ops.linear_batched(ws["KV"][:, C:], w["d_o_w"], ws["U"], M2, C, hdd, Hd, hdd, hdd, 2 * C, ldx=2 * C, ldw=C, ymap=(2, R, 0))
        ops.cdm_decoder_point(x, cond.xyz, w["dd_w"], w["dd_b"], w["d_qn_g"], w["d_qn_b"], ws["KF"], C + 4, ws["U"], w["d_o_b"],
                              w["d_m_g"], w["d_m_b"], ws["H1"], ws["HN"], B, N, cx, hn2=ws["HN2"])
        if self.gemm == "tc":  # the one dense per-point layer: 256 -> 256 GELU on the tensor cores
            ops.linear_tc(ws["HN2"], w["d_m1_w2"], B * N, C, C, y=ws["Gm"], bias=w["d_m1_b"], act="gelu")
        else:
            ops.linear(ws["HN"], w["d_m1_w"], ws["Gm"], B * N, C, C, bias=w["d_m1_b"], act="gelu")
        ops.linear_skinny(ws["H1"], C, ws["Gm"], C, w["head_w"], w["head_b"], out, B * N, self.out_dim)
        return out
