"""am_linear_ln_tc (csrc/gemm_ln_tc.cu): GEMM + bias + residual + LayerNorm in one tcgen05 kernel vs torch fp64, and vs the unfused
am_linear_tc + am_layernorm pair it replaces (out_proj + norm1, linear2 + norm2 of the CMDM trunk, models/cmdm.py:66-77)."""
import pytest
import torch

from amb200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("M,K", [(10432, 512), (10432, 1024), (300, 512), (257, 64), (128, 1024)])
def test_linear_ln_tc_vs_fp64_and_unfused(M, K):
    N = 512
    g = torch.Generator().manual_seed(M + K)
    x, w, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5, torch.randn(N, generator=g)
    r = torch.randn(M, N, generator=g)
    gam, bet = 1 + 0.1 * torch.randn(N, generator=g), 0.1 * torch.randn(N, generator=g)
    a2, w2, r2 = ops.split_bf16(x.to(DEV), M, K), ops.split_bf16(w.to(DEV), N, K), ops.split_bf16(r.to(DEV), M, N)
    y2 = torch.full((M, 2 * N), float("nan"), dtype=torch.bfloat16, device=DEV)
    ops.linear_ln_tc(a2, w2, M, N, K, b.to(DEV), r2, gam.to(DEV), bet.to(DEV), 1e-5, y2)
    got = (y2[:, :N].float() + y2[:, N:].float()).cpu().double()
    rr = (r2[:, :N].float() + r2[:, N:].float()).cpu().double()  # the residual the kernel sees (16 significant bits)
    pre = x.double() @ w.double().T + b.double() + rr
    ref = torch.nn.functional.layer_norm(pre, (N,), gam.double(), bet.double(), 1e-5)
    err = (got - ref).abs().max().item()
    assert err < 1e-4, err
    # the unfused pair
    tmp = torch.empty(M, N, device=DEV)
    ops.linear_tc(a2, w2, M, N, K, y=tmp, bias=b.to(DEV), residual_split=r2)
    y2u = torch.zeros(M, 2 * N, dtype=torch.bfloat16, device=DEV)
    ops.layernorm(tmp, gam.to(DEV), bet.to(DEV), None, M, N, eps=1e-5, y2=y2u)
    un = (y2u[:, :N].float() + y2u[:, N:].float()).cpu().double()
    assert (got - un).abs().max().item() < 2e-4  # both round to bf16 pairs (16 significant bits) from slightly different fp32 values
