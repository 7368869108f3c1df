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


@pytest.mark.parametrize("M,K,seg,q0", [(10432, 512, 326, 130), (10432, 1024, 0, 0), (6272, 512, 0, 0), (978, 512, 326, 130)])
def test_layernorm_overlapped_with_its_gemm_is_bit_identical(M, K, seg, q0):
    """am_linear_tc_set_rowflags + am_layernorm_flags (the LayerNorm consumes 128-row blocks while the GEMM is still running, handshake
    through per-block completion counters) == am_linear_tc followed by am_layernorm(_win), bit for bit, over repeated launches (the
    counters are reset by the consumer) and back to back with other kernels in between."""
    N = 512
    g = torch.Generator().manual_seed(M + K + 7)
    x, w, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5, torch.randn(N, generator=g)
    r = torch.randn(M, N, generator=g)
    gam, bet = (1 + 0.1 * torch.randn(N, generator=g)).to(DEV), (0.1 * torch.randn(N, generator=g)).to(DEV)
    a2, w2, r2 = ops.split_bf16(x.to(DEV), M, K), ops.split_bf16(w.to(DEV), N, K), ops.split_bf16(r.to(DEV), M, N)
    bd = b.to(DEV)
    nwin = (M // seg) * (seg - q0) if seg else 0
    tmp = torch.empty(M, N, device=DEV)
    ref = torch.zeros(M, 2 * N, dtype=torch.bfloat16, device=DEV)
    refw = torch.zeros(max(nwin, 1), 2 * N, dtype=torch.bfloat16, device=DEV)
    ops.linear_tc(a2, w2, M, N, K, y=tmp, bias=bd, residual_split=r2)
    if seg:
        ops.layernorm(tmp, gam, bet, None, M, N, eps=1e-5, y2=ref, y2_win=refw, seg=seg, seg_q0=q0)
    else:
        ops.layernorm(tmp, gam, bet, None, M, N, eps=1e-5, y2=ref)
    flags = torch.zeros((M + 127) // 128 + 1, dtype=torch.int32, device=DEV)
    tmp2 = torch.empty(M, N, device=DEV)
    for it in range(12):
        got = torch.full((M, 2 * N), float("nan"), dtype=torch.bfloat16, device=DEV)
        gotw = torch.full((max(nwin, 1), 2 * N), float("nan"), dtype=torch.bfloat16, device=DEV)
        tmp2.fill_(float("nan"))
        ops.linear_tc(a2, w2, M, N, K, y=tmp2, bias=bd, residual_split=r2, rowflags=flags)
        ops.layernorm_flags(tmp2, gam, bet, M, N, got, flags, 4 * N, eps=1e-5, **(dict(y2_win=gotw, seg=seg, seg_q0=q0) if seg else {}))
        if it % 3 == 0:   # another GEMM right behind (the next kernel of the trunk): it must see the LayerNorm's output complete
            y2n = torch.zeros(M, 2 * N, dtype=torch.bfloat16, device=DEV)
            ops.linear_tc(got, ops.split_bf16(torch.eye(N, device=DEV), N, N), M, N, N, y2=y2n, Np2=N)
        assert torch.equal(got.view(torch.int16), ref.view(torch.int16)), it
        if seg:
            assert torch.equal(gotw.view(torch.int16), refw.view(torch.int16)), it
        assert int(flags[: (M + 127) // 128].abs().sum()) == 0   # counters reset by the consumer


@pytest.mark.parametrize("M,K", [(10432, 512), (6272, 1024), (300, 512)])
def test_residual_added_in_layernorm_is_bit_identical(M, K):
    """am_linear_tc without a residual + am_layernorm_win(Rsplit=...) == am_linear_tc(residual_split=...) + am_layernorm: the same fp32
    sum (acc + bias) + (hi + lo), formed in the LayerNorm's load phase instead of the GEMM epilogue."""
    N = 512
    g = torch.Generator().manual_seed(M + K + 11)
    x, w, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5, torch.randn(N, generator=g)
    r = torch.randn(M, N, generator=g)
    gam, bet = (1 + 0.1 * torch.randn(N, generator=g)).to(DEV), (0.1 * torch.randn(N, generator=g)).to(DEV)
    a2, w2, r2 = ops.split_bf16(x.to(DEV), M, K), ops.split_bf16(w.to(DEV), N, K), ops.split_bf16(r.to(DEV), M, N)
    bd = b.to(DEV)
    tmp = torch.empty(M, N, device=DEV)
    ref = torch.zeros(M, 2 * N, dtype=torch.bfloat16, device=DEV)
    ops.linear_tc(a2, w2, M, N, K, y=tmp, bias=bd, residual_split=r2)
    ops.layernorm(tmp, gam, bet, None, M, N, eps=1e-5, y2=ref)
    tmp2 = torch.empty(M, N, device=DEV)
    got = torch.zeros(M, 2 * N, dtype=torch.bfloat16, device=DEV)
    ops.linear_tc(a2, w2, M, N, K, y=tmp2, bias=bd)
    ops.layernorm(tmp2, gam, bet, None, M, N, eps=1e-5, y2=got, residual_split=r2)
    assert torch.equal(got.view(torch.int16), ref.view(torch.int16))
