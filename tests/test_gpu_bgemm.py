"""am_gemm_f32 on batched shapes (csrc/bgemm_tc.cu: tcgen05 with on-the-fly bf16 hi|lo conversion) — the six attention products of the
training step (amb200/autograd_ops.py AttentionFn) as NT / NN / TN forms with the strides the autograd functions pass — vs torch fp64."""
import pytest
import torch

from amb200 import autograd_ops as A

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("B,H,S,hd", [(2, 8, 326, 64), (3, 8, 70, 64), (1, 8, 129, 64)])
def test_batched_attention_products_vs_fp64(B, H, S, hd):
    g = torch.Generator().manual_seed(S)
    D = H * hd
    D3 = 3 * D
    qkv = torch.randn(B, S, D3, generator=g)
    P = torch.rand(B * H, S, S, generator=g)
    dO = torch.randn(B, S, D, generator=g)
    q, k, v = (t.reshape(B, S, H, hd).permute(0, 2, 1, 3).reshape(B * H, S, hd).double() for t in qkv.split(D, -1))
    dOh = dO.reshape(B, S, H, hd).permute(0, 2, 1, 3).reshape(B * H, S, hd).double()
    qkv_d, P_d, dO_d = qkv.to(DEV), P.to(DEV), dO.to(DEV)
    sq, so, sp = (S * D3, hd), (S * D, hd), (H * S * S, S * S)

    def rel(got, ref):
        return ((got.cpu().double() - ref).abs().max() / ref.abs().max()).item()
    # NT: S = Q K^T
    out = torch.empty(B * H, S, S, device=DEV)
    A.gemm(qkv_d, qkv_d[:, :, D:], out, S, S, hd, transB=True, lda=D3, ldb=D3, ldc=S, batch=B * H, bdiv=H, sA=sq, sB=sq, sC=sp)
    assert rel(out, q @ k.transpose(1, 2)) < 2e-5
    # NN: O = P V  (strided output [B,S,D])
    o = torch.zeros(B, S, D, device=DEV)
    A.gemm(P_d, qkv_d[:, :, 2 * D:], o, S, hd, S, lda=S, ldb=D3, ldc=D, batch=B * H, bdiv=H, sA=sp, sB=sq, sC=so)
    ref = (P.double() @ v).reshape(B, H, S, hd).permute(0, 2, 1, 3).reshape(B, S, D)
    assert rel(o, ref) < 2e-5
    # TN: dV = P^T dO  (strided output inside dqkv)
    dqkv = torch.zeros(B, S, D3, device=DEV)
    A.gemm(P_d, dO_d, dqkv[:, :, 2 * D:], S, hd, S, transA=True, lda=S, ldb=D, ldc=D3, batch=B * H, bdiv=H, sA=sp, sB=so, sC=sq)
    ref = (P.double().transpose(1, 2) @ dOh).reshape(B, H, S, hd).permute(0, 2, 1, 3).reshape(B, S, D)
    assert rel(dqkv[:, :, 2 * D:], ref) < 2e-5
    assert float(dqkv[:, :, :2 * D].abs().max()) == 0.0  # nothing written outside the V third
    # NT with a strided A: dP = dO V^T, and alpha / beta
    dP = torch.ones(B * H, S, S, device=DEV)
    A.gemm(dO_d, qkv_d[:, :, 2 * D:], dP, S, S, hd, transB=True, alpha=0.5, beta=2.0, lda=D, ldb=D3, ldc=S, batch=B * H, bdiv=H, sA=so, sB=sq, sC=sp)
    assert rel(dP, 0.5 * (dOh @ v.transpose(1, 2)) + 2.0) < 2e-5
