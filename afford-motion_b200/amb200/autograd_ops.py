"""torch.autograd Functions whose forward AND backward are libamb200 kernels — the training path of the CMDM denoiser
(SURVEY §8 a10/a28/a29).  torch supplies only the tape, parameter storage and data movement (cat / slice / view);
every arithmetic op below runs a kernel from csrc/train_kernels.cu (fp32 SIMT; tensor-core versions are next)."""
from typing import Optional

import torch
import torch.distributed as dist
from torch.autograd import Function

from . import lib as _l
from . import ops as _ops

_L = None


def L():
    global _L
    if _L is None:
        _L = _l.load()
    return _L


def _st():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


def gemm(A, B, C, M, N, K, transA=False, transB=False, alpha=1.0, beta=0.0, lda=None, ldb=None, ldc=None, batch=1, bdiv=1,
         sA=(0, 0), sB=(0, 0), sC=(0, 0)):
    lda = (M if transA else K) if lda is None else lda
    ldb = (K if transB else N) if ldb is None else ldb
    ldc = N if ldc is None else ldc
    _l.check(L().am_gemm_f32(int(transA), int(transB), M, N, K, float(alpha), _p(A), lda, _p(B), ldb, float(beta), _p(C), ldc, batch, bdiv,
                             sA[0], sA[1], sB[0], sB[1], sC[0], sC[1], _st()), "am_gemm_f32")
    return C


def _tsplit(x2, M, K):
    """bf16 (hi|lo) of x2^T: [K, 2*Mp]."""
    Mp = _ops.pad32(M)
    out = torch.empty(K, 2 * Mp, dtype=torch.bfloat16, device=x2.device)
    _l.check(L().am_transpose_split_bf16(_p(x2), x2.stride(0), _p(out), Mp, M, K, _st()), "am_transpose_split_bf16")
    return out, Mp


def _use_tc(M, N, K):
    """Large trunk layers go to the tcgen05 GEMM (3-term bf16 split, fp32 accumulate); small / odd shapes stay SIMT."""
    return _TC_TRAIN and M >= 2048 and N >= 32 and K >= 32 and N % 4 == 0 and K % 4 == 0


import os as _os
_TC_TRAIN = _os.environ.get("AMB200_TRAIN_GEMM", "tc") == "tc"


class LinearFn(Function):
    """y = x W^T (+ b) on the last dim; x [..., K], W [N, K]."""

    @staticmethod
    def forward(ctx, x, w, b, tc=False):
        x2 = _c(x).view(-1, x.shape[-1])
        w = _c(w)
        M, K = x2.shape
        N = w.shape[0]
        y = torch.empty(M, N, device=x.device)
        ctx.tc = tc = bool(tc) and _use_tc(M, N, K)
        if tc:
            _ops.linear_tc(_ops.split_bf16(x2, M, K), _ops.split_bf16(w, N, K), M, N, _ops.pad32(K), y=y, bias=None if b is None else _c(b))
        else:
            _ops.linear(x2, w, y, M, N, K, bias=None if b is None else _c(b))  # fp32 SIMT GEMM with the bias fused in the epilogue
        ctx.save_for_backward(x2, w)
        ctx.has_bias = b is not None
        ctx.xshape = x.shape
        return y.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        x2, w = ctx.saved_tensors
        M, K = x2.shape
        N = w.shape[0]
        dy2 = _c(dy).view(M, N)
        dx = dw = db = None
        tc = ctx.tc
        if ctx.needs_input_grad[0]:
            dx = torch.empty(M, K, device=dy.device)
            if tc:  # dX = dY W : A = dY (K-major along N), B = W^T (rows = K_in, K-major along N)
                wt2, Np = _tsplit(w, N, K)
                _ops.linear_tc(_ops.split_bf16(dy2, M, N), wt2, M, K, Np, y=dx)
            else:
                gemm(dy2, w, dx, M, K, N)
            dx = dx.view(ctx.xshape)
        if ctx.needs_input_grad[1]:
            dw = torch.empty(N, K, device=dy.device)
            if tc:  # dW = dY^T X : both operands K-major along the M batch rows
                dyt2, Mp = _tsplit(dy2, M, N)
                xt2, _ = _tsplit(x2, M, K)
                _ops.linear_tc(dyt2, xt2, N, K, Mp, y=dw)
            else:
                gemm(dy2, x2, dw, N, K, M, transA=True)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = torch.empty(N, device=dy.device)
            _l.check(L().am_colsum_f32(_p(dy2), N, _p(db), M, N, 0.0, _st()), "am_colsum_f32")
        return dx, dw, db, None


def _unary(fwd, bwd, name):
    class Fn(Function):
        @staticmethod
        def forward(ctx, x):
            x = _c(x)
            y = torch.empty_like(x)
            _l.check(getattr(L(), fwd)(_p(x), _p(y), x.numel(), _st()), fwd)
            ctx.save_for_backward(x)
            return y

        @staticmethod
        def backward(ctx, dy):
            (x,) = ctx.saved_tensors
            dy = _c(dy)
            dx = torch.empty_like(x)
            _l.check(getattr(L(), bwd)(_p(dy), _p(x), _p(dx), x.numel(), _st()), bwd)
            return dx
    Fn.__name__ = name
    return Fn


GeluFn = _unary("am_gelu_fwd", "am_gelu_bwd", "GeluFn")
SiluFn = _unary("am_silu_fwd", "am_silu_bwd", "SiluFn")


class AddFn(Function):
    """y = a + b (optionally ReLU), same shapes."""

    @staticmethod
    def forward(ctx, a, b, relu):
        a, b = _c(a), _c(b)
        y = torch.empty_like(a)
        _l.check(L().am_add_f32(_p(a), _p(b), _p(y), a.numel(), int(relu), _st()), "am_add_f32")
        ctx.relu = relu
        if relu:
            ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = _c(dy)
        if ctx.relu:
            (y,) = ctx.saved_tensors
            d = torch.empty_like(dy)
            _l.check(L().am_relu_bwd(_p(dy), _p(y), _p(d), dy.numel(), _st()), "am_relu_bwd")
            dy = d
        return dy, dy, None


class DropoutFn(Function):
    @staticmethod
    def forward(ctx, x, p, seed, site):
        if p <= 0.0:
            ctx.p = 0.0
            return x
        x = _c(x)
        y = torch.empty_like(x)
        _l.check(L().am_dropout(_p(x), _p(y), x.numel(), float(p), seed, site, _st()), "am_dropout")
        ctx.p, ctx.seed, ctx.site = p, seed, site
        return y

    @staticmethod
    def backward(ctx, dy):
        if ctx.p <= 0.0:
            return dy, None, None, None
        dy = _c(dy)
        dx = torch.empty_like(dy)
        _l.check(L().am_dropout(_p(dy), _p(dx), dy.numel(), float(ctx.p), ctx.seed, ctx.site, _st()), "am_dropout")
        return dx, None, None, None


class LayerNormFn(Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        x2 = _c(x).view(-1, x.shape[-1])
        M, D = x2.shape
        y = torch.empty_like(x2)
        _l.check(L().am_layernorm(_p(x2), D, None, D, _p(gamma), _p(beta), _p(y), D, M, D, float(eps), None, 0, _st()), "am_layernorm")
        ctx.save_for_backward(x2, gamma)
        ctx.eps, ctx.shape = eps, x.shape
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, gamma = ctx.saved_tensors
        M, D = x2.shape
        dy2 = _c(dy).view(M, D)
        dx = torch.empty_like(x2)
        dg, db = torch.zeros(D, device=dy.device), torch.zeros(D, device=dy.device)
        _l.check(L().am_layernorm_bwd(_p(dy2), _p(x2), None, _p(gamma), _p(dx), _p(dg), _p(db), M, D, float(ctx.eps), _st()), "am_layernorm_bwd")
        return dx.view(ctx.shape), dg, db, None


class AttentionFn(Function):
    """softmax(q k^T * scale + key mask) (dropout) v on packed qkv [B,S,3*H*hd] -> [B,S,H*hd]; the probabilities are kept
    for the backward pass (training only; sampling uses the tcgen05 kernel)."""

    @staticmethod
    def forward(ctx, qkv, key_pad_u8, H, p_drop, seed, site):
        qkv = _c(qkv)
        B, S, D3 = qkv.shape
        D = D3 // 3
        hd = D // H
        scale = hd ** -0.5
        P = torch.empty(B * H, S, S, device=qkv.device)
        sq = (S * D3, hd)
        gemm(qkv, qkv[:, :, D:], P, S, S, hd, transB=True, lda=D3, ldb=D3, ldc=S, batch=B * H, bdiv=H, sA=sq, sB=sq, sC=(H * S * S, S * S))
        _l.check(L().am_softmax_rows_fwd(_p(P), _p(key_pad_u8), B * H * S, S, H * S, float(scale), _st()), "am_softmax_rows_fwd")
        Pd = P
        if p_drop > 0:
            Pd = torch.empty_like(P)
            _l.check(L().am_dropout(_p(P), _p(Pd), P.numel(), float(p_drop), seed, site, _st()), "am_dropout")
        out = torch.empty(B, S, D, device=qkv.device)
        gemm(Pd, qkv[:, :, 2 * D:], out, S, hd, S, lda=S, ldb=D3, ldc=D, batch=B * H, bdiv=H, sA=(H * S * S, S * S), sB=sq, sC=(S * D, hd))
        ctx.save_for_backward(qkv, P, Pd)
        ctx.meta = (B, S, D, H, hd, scale, p_drop, seed, site)
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv, P, Pd = ctx.saved_tensors
        B, S, D, H, hd, scale, p_drop, seed, site = ctx.meta
        D3 = 3 * D
        dout = _c(dout)
        dqkv = torch.zeros_like(qkv)
        sq, so, sp = (S * D3, hd), (S * D, hd), (H * S * S, S * S)
        dP = torch.empty_like(P)
        gemm(dout, qkv[:, :, 2 * D:], dP, S, S, hd, transB=True, lda=D, ldb=D3, ldc=S, batch=B * H, bdiv=H, sA=so, sB=sq, sC=sp)  # dO V^T
        gemm(Pd, dout, dqkv[:, :, 2 * D:], S, hd, S, transA=True, lda=S, ldb=D, ldc=D3, batch=B * H, bdiv=H, sA=sp, sB=so, sC=sq)  # dV = Pd^T dO
        if p_drop > 0:
            d2 = torch.empty_like(dP)
            _l.check(L().am_dropout(_p(dP), _p(d2), dP.numel(), float(p_drop), seed, site, _st()), "am_dropout")
            dP = d2
        _l.check(L().am_softmax_rows_bwd(_p(dP), _p(P), B * H * S, S, float(scale), _st()), "am_softmax_rows_bwd")  # dS (scaled)
        gemm(dP, qkv[:, :, D:], dqkv, S, hd, S, lda=S, ldb=D3, ldc=D3, batch=B * H, bdiv=H, sA=sp, sB=sq, sC=sq)  # dQ = dS K
        gemm(dP, qkv, dqkv[:, :, D:], S, hd, S, transA=True, lda=S, ldb=D3, ldc=D3, batch=B * H, bdiv=H, sA=sp, sB=sq, sC=sq)  # dK = dS^T Q
        return dqkv, None, None, None, None, None


class BatchNormTrainFn(Function):
    """BatchNorm1d (training: batch statistics over the rows of x [M,C]) with optional fused ReLU.  Under
    torch.distributed + SyncBatchNorm the (sum, sumsq) / (dbeta, dgamma) accumulators are all-reduced (train_ddp.py:63)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, bn, relu):
        x = _c(x)
        M, C = x.shape
        dev = x.device
        sync = isinstance(bn, torch.nn.SyncBatchNorm) and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        acc = torch.zeros(2 * C + 1, dtype=torch.float64, device=dev)
        mean, invstd, var = (torch.empty(C, device=dev) for _ in range(3))
        cnt = None
        _l.check(L().am_bn_train_stats(_p(x), M, C, float(bn.eps), _p(acc), _p(mean), _p(invstd), _p(var), _st()), "am_bn_train_stats")
        if sync:
            # global statistics: (sum, sumsq, row count) over all ranks in ONE all-reduce; the count stays on the device — a
            # host read here (.item()) would drain the GPU once per BatchNorm layer, 28 times per CMDM training step
            acc[2 * C:].fill_(float(M))
            dist.all_reduce(acc)
            cnt = acc[2 * C:]
        track = bn.track_running_stats and bn.running_mean is not None
        if sync or track:
            # finalise (mean / invstd / var from the global sums) + running-statistics update in ONE launch (torch semantics:
            # unbiased variance, momentum); under SyncBatchNorm this replaces ~8 tiny launches behind every all-reduce
            mom = bn.momentum if bn.momentum is not None else 0.1
            with torch.no_grad():
                _l.check(L().am_bn_finalize_running(_p(acc), _p(cnt) if sync else None, M, C, float(bn.eps), _p(mean), _p(invstd), _p(var),
                                                    _p(bn.running_mean) if track else None, _p(bn.running_var) if track else None, float(mom),
                                                    _p(bn.num_batches_tracked) if track else None, _st()), "am_bn_finalize_running")
                if track:  # the kernel wrote the buffers behind torch's back: bump their version counters (amb200.pack.params_version)
                    torch.autograd.graph.increment_version([bn.running_mean, bn.running_var, bn.num_batches_tracked])
        y = torch.empty_like(x)
        _l.check(L().am_bn_apply(_p(x), _p(mean), _p(invstd), _p(gamma), _p(beta), _p(y), M, C, int(relu), _st()), "am_bn_apply")
        ctx.save_for_backward(x, y if relu else x, mean, invstd, gamma)
        ctx.relu, ctx.sync, ctx.cnt = relu, sync, cnt
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, mean, invstd, gamma = ctx.saved_tensors
        M, C = x.shape
        dy = _c(dy)
        acc = torch.zeros(2 * C, dtype=torch.float64, device=dy.device)
        dx = torch.empty_like(x)
        dg, db = torch.zeros(C, device=dy.device), torch.zeros(C, device=dy.device)
        if not ctx.sync:
            _l.check(L().am_bn_bwd(_p(dy), _p(x), _p(y), _p(mean), _p(invstd), _p(gamma), _p(acc), _p(dx), _p(dg), _p(db), M, C, int(ctx.relu), _st()),
                     "am_bn_bwd")
        else:
            # SyncBatchNorm: local (sum dy', sum dy' xhat) -> parameter gradients are the LOCAL sums (DDP averages them over ranks,
            # same as torch.nn.SyncBatchNorm), the input gradient uses the GLOBAL sums and the global row count
            _l.check(L().am_bn_bwd_reduce(_p(dy), _p(x), _p(y), _p(mean), _p(invstd), _p(acc), M, C, int(ctx.relu), _st()), "am_bn_bwd_reduce")
            db, dg = acc[:C].float(), acc[C:].float()
            dist.all_reduce(acc)
            # the kernel divides by its integer row count: scale the global sums by M_local / M_global on the device instead
            # of reading the global count back to the host
            acc = acc * (float(M) / ctx.cnt)
            scratch_g, scratch_b = torch.zeros(C, device=dy.device), torch.zeros(C, device=dy.device)
            _l.check(L().am_bn_bwd_apply(_p(dy), _p(x), _p(y), _p(mean), _p(invstd), _p(gamma), _p(acc), _p(dx), _p(scratch_g), _p(scratch_b), M,
                                         M, C, int(ctx.relu), _st()), "am_bn_bwd_apply")
        return dx, dg, db, None, None


class GroupCatFn(Function):
    """G = cat(rel, x[idx]) for TransitionDown (pointtransformer.py:63); gradient flows to x by scatter-add."""

    @staticmethod
    def forward(ctx, rel, x, idx):
        x = _c(x)
        mk, c = idx.numel(), x.shape[1]
        G = torch.empty(mk, 3 + c, device=x.device)
        _l.check(L().am_group_cat(_p(rel), _p(x), _p(idx), _p(G), mk, c, _st()), "am_group_cat")
        ctx.save_for_backward(idx)
        ctx.n, ctx.c = x.shape[0], c
        return G

    @staticmethod
    def backward(ctx, dG):
        (idx,) = ctx.saved_tensors
        dG = _c(dG)
        dx = torch.zeros(ctx.n, ctx.c, device=dG.device)
        _l.check(L().am_scatter_add_rows(_p(dG), 3 + ctx.c, 3, _p(idx), _p(dx), idx.numel(), ctx.c, _st()), "am_scatter_add_rows")
        return None, dx, None


class MaxPoolKFn(Function):
    @staticmethod
    def forward(ctx, Z, m, k):
        Z = _c(Z)
        c = Z.shape[1]
        out = torch.empty(m, c, device=Z.device)
        arg = torch.empty(m, c, dtype=torch.int32, device=Z.device)
        _l.check(L().am_maxpool_k_fwd(_p(Z), _p(out), _p(arg), m, k, c, _st()), "am_maxpool_k_fwd")
        ctx.save_for_backward(arg)
        ctx.meta = (m, k, c)
        return out

    @staticmethod
    def backward(ctx, dout):
        (arg,) = ctx.saved_tensors
        m, k, c = ctx.meta
        dZ = torch.empty(m * k, c, device=dout.device)
        _l.check(L().am_maxpool_k_bwd(_p(_c(dout)), _p(arg), _p(dZ), m, k, c, _st()), "am_maxpool_k_bwd")
        return dZ, None, None


class PtWFn(Function):
    """w = k[idx] - q + pr  (pointtransformer.py:33)."""

    @staticmethod
    def forward(ctx, qkv, pr, idx, k):
        qkv, pr = _c(qkv), _c(pr)
        n, c = qkv.shape[0], qkv.shape[1] // 3
        w = torch.empty(n * k, c, device=qkv.device)
        _l.check(L().am_pt_w_fwd(_p(qkv), _p(idx), _p(pr), _p(w), n, k, c, _st()), "am_pt_w_fwd")
        ctx.save_for_backward(idx)
        ctx.meta = (n, k, c)
        return w

    @staticmethod
    def backward(ctx, dw):
        (idx,) = ctx.saved_tensors
        n, k, c = ctx.meta
        dw = _c(dw)
        dqkv = torch.zeros(n, 3 * c, device=dw.device)
        dpr = torch.zeros(n * k, c, device=dw.device)
        _l.check(L().am_pt_w_bwd(_p(dw), _p(idx), _p(dqkv), _p(dpr), n, k, c, _st()), "am_pt_w_bwd")
        return dqkv, dpr, None, None


class SoftmaxKFn(Function):
    @staticmethod
    def forward(ctx, w, n, k):
        w = _c(w).clone()
        c8 = w.shape[1]
        _l.check(L().am_softmax_k_fwd(_p(w), n, k, c8, _st()), "am_softmax_k_fwd")
        ctx.save_for_backward(w)
        ctx.meta = (n, k, c8)
        return w

    @staticmethod
    def backward(ctx, dw):
        (w,) = ctx.saved_tensors
        n, k, c8 = ctx.meta
        d = _c(dw).clone()
        _l.check(L().am_softmax_k_bwd(_p(d), _p(w), n, k, c8, _st()), "am_softmax_k_bwd")
        return d, None, None


class PtAggFn(Function):
    """out[i, ch] = sum_j (v[idx[i,j], ch] + pr[i,j,ch]) * ws[i,j, ch % c8]  (pointtransformer.py:36-37)."""

    @staticmethod
    def forward(ctx, qkv, pr, ws, idx, k):
        qkv, pr, ws = _c(qkv), _c(pr), _c(ws)
        n, c = qkv.shape[0], qkv.shape[1] // 3
        out = torch.empty(n, c, device=qkv.device)
        _l.check(L().am_pt_agg_fwd(_p(qkv), _p(idx), _p(pr), _p(ws), _p(out), n, k, c, _st()), "am_pt_agg_fwd")
        ctx.save_for_backward(qkv, pr, ws, idx)
        ctx.meta = (n, k, c)
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv, pr, ws, idx = ctx.saved_tensors
        n, k, c = ctx.meta
        dqkv = torch.zeros_like(qkv)
        dpr = torch.empty_like(pr)
        dws = torch.empty_like(ws)
        _l.check(L().am_pt_agg_bwd(_p(_c(dout)), _p(qkv), _p(idx), _p(pr), _p(ws), _p(dqkv), _p(dpr), _p(dws), n, k, c, _st()), "am_pt_agg_bwd")
        return dqkv, dpr, dws, None, None


class MaskedMSEFn(Function):
    """loss[b] of gaussian_diffusion.py:815-818 (differentiable w.r.t. the model output)."""

    @staticmethod
    def forward(ctx, pred, target, mask_u8):
        pred, target = _c(pred), _c(target)
        B, T, D = pred.shape
        loss = torch.empty(B, device=pred.device)
        _l.check(L().am_masked_mse(_p(target), _p(pred), _p(mask_u8), _p(loss), B, T, D, _st()), "am_masked_mse")
        ctx.save_for_backward(pred, target, mask_u8)
        return loss

    @staticmethod
    def backward(ctx, gloss):
        pred, target, mask = ctx.saved_tensors
        B, T, D = pred.shape
        d = torch.empty_like(pred)
        _l.check(L().am_masked_mse_bwd(_p(target), _p(pred), _p(mask), _p(_c(gloss)), _p(d), B, T, D, _st()), "am_masked_mse_bwd")
        return d, None, None


# ---- functional helpers
def linear(x, w, b=None, tc=False):
    """tc=True: large layers run on the tcgen05 GEMM (3-term bf16 split, ~1e-5 relative); used by the LayerNorm trunks.  The
    BatchNorm scene encoder keeps fp32 SIMT GEMMs: its deep stages normalise over a handful of points and amplify GEMM noise."""
    return LinearFn.apply(x, w, b, tc)


def bn_train(x, bn, relu=False):
    return BatchNormTrainFn.apply(x, bn.weight, bn.bias, bn, relu)


class MHAFn(Function):
    """General multi-head attention core on separate q [B,Sq,C], k [B,Sk,C], v [B,Sk,Cv] (no mask):
    softmax(q_h k_h^T * scale) (dropout) v_h, heads split along the channel dim — the Perceiver cross/self attention of
    models/modules.py:324-381.  Batched strided GEMMs + row softmax kernels; probabilities kept for backward."""

    @staticmethod
    def forward(ctx, q, k, v, H, scale, p_drop, seed, site):
        q, k, v = _c(q), _c(k), _c(v)
        B, Sq, C = q.shape
        Sk, Cv = k.shape[1], v.shape[2]
        c, cv = C // H, Cv // H
        P = torch.empty(B * H, Sq, Sk, device=q.device)
        sP = (H * Sq * Sk, Sq * Sk)
        gemm(q, k, P, Sq, Sk, c, transB=True, lda=C, ldb=C, ldc=Sk, batch=B * H, bdiv=H, sA=(Sq * C, c), sB=(Sk * C, c), sC=sP)
        _l.check(L().am_softmax_rows_fwd(_p(P), None, B * H * Sq, Sk, H * Sq, float(scale), _st()), "am_softmax_rows_fwd")
        Pd = P
        if p_drop > 0:
            Pd = torch.empty_like(P)
            _l.check(L().am_dropout(_p(P), _p(Pd), P.numel(), float(p_drop), seed, site, _st()), "am_dropout")
        out = torch.empty(B, Sq, Cv, device=q.device)
        gemm(Pd, v, out, Sq, cv, Sk, lda=Sk, ldb=Cv, ldc=Cv, batch=B * H, bdiv=H, sA=sP, sB=(Sk * Cv, cv), sC=(Sq * Cv, cv))
        ctx.save_for_backward(q, k, v, P, Pd)
        ctx.meta = (B, Sq, Sk, C, Cv, H, c, cv, scale, p_drop, seed, site)
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, P, Pd = ctx.saved_tensors
        B, Sq, Sk, C, Cv, H, c, cv, scale, p_drop, seed, site = ctx.meta
        dout = _c(dout)
        sP = (H * Sq * Sk, Sq * Sk)
        sq, sk, sv, so = (Sq * C, c), (Sk * C, c), (Sk * Cv, cv), (Sq * Cv, cv)
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        dP = torch.empty_like(P)
        gemm(dout, v, dP, Sq, Sk, cv, transB=True, lda=Cv, ldb=Cv, ldc=Sk, batch=B * H, bdiv=H, sA=so, sB=sv, sC=sP)       # dO V^T
        gemm(Pd, dout, dv, Sk, cv, Sq, transA=True, lda=Sk, ldb=Cv, ldc=Cv, batch=B * H, bdiv=H, sA=sP, sB=so, sC=sv)      # dV = Pd^T dO
        if p_drop > 0:
            d2 = torch.empty_like(dP)
            _l.check(L().am_dropout(_p(dP), _p(d2), dP.numel(), float(p_drop), seed, site, _st()), "am_dropout")
            dP = d2
        _l.check(L().am_softmax_rows_bwd(_p(dP), _p(P), B * H * Sq, Sk, float(scale), _st()), "am_softmax_rows_bwd")
        gemm(dP, k, dq, Sq, c, Sk, lda=Sk, ldb=C, ldc=C, batch=B * H, bdiv=H, sA=sP, sB=sk, sC=sq)                          # dQ = dS K
        gemm(dP, q, dk, Sk, c, Sq, transA=True, lda=Sk, ldb=C, ldc=C, batch=B * H, bdiv=H, sA=sP, sB=sq, sC=sk)            # dK = dS^T Q
        return dq, dk, dv, None, None, None, None, None
