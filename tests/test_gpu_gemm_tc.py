"""GPU: the tcgen05/TMA GEMM (impl=2, bf16 in, fp32 accumulate) against an fp64 product of
the same bf16-rounded operands, for the three contractions of a linear layer, including
ragged tails in every dimension and strided column-block views."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture()
def tc_ops(cuda_device):
    from statecatcher_b200 import ops
    old = ops.GEMM_IMPL
    ops.GEMM_IMPL = 2                      # force tcgen05: unsupported shapes raise instead of using SIMT
    yield ops
    ops.GEMM_IMPL = old


def _rel(got, want):
    got = got.double().cpu()
    return ((got - want).abs().max() / want.abs().max().clamp_min(1e-6)).item()


SHAPES = [(128, 256, 64), (256, 512, 128), (1000, 1280, 256), (192, 1024, 80), (4096, 5120, 1024),
          (130, 264, 72), (77, 64, 64), (3001, 1024, 1024)]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_tc_fwd_dgrad_wgrad(tc_ops, M, N, K):
    ops = tc_ops
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g, device="cuda").bfloat16()
    w = (torch.randn(N, K, generator=g, device="cuda") / K ** 0.5).bfloat16()
    dy = torch.randn(M, N, generator=g, device="cuda").bfloat16()
    bias = torch.randn(N, generator=g, device="cuda")
    A, W, DY = a.double().cpu(), w.double().cpu(), dy.double().cpu()
    y = ops.gemm_fwd(a, w, bias)                                   # bf16 out
    assert y.dtype == torch.bfloat16 and _rel(y, A @ W.T + bias.double().cpu()) < 1e-2
    y32 = ops.gemm_fwd(a, w, bias, out_dtype=torch.float32)        # fp32 out: only accumulation error
    assert _rel(y32, A @ W.T + bias.double().cpu()) < 2e-5
    y32n = ops.gemm_fwd(a, w, None, out_dtype=torch.float32)
    assert _rel(y32n, A @ W.T) < 2e-5
    da = ops.gemm_dgrad(dy, w, out_dtype=torch.float32)
    assert _rel(da, DY @ W) < 2e-5
    dab = ops.gemm_dgrad(dy, w)
    assert dab.dtype == torch.bfloat16 and _rel(dab, DY @ W) < 1e-2
    if M >= 64:
        dw = ops.gemm_wgrad(dy, a)
        assert dw.dtype == torch.float32 and _rel(dw, DY.T @ A) < 2e-5
        base = torch.randn(N, K, generator=g, device="cuda")
        dw2 = ops.gemm_wgrad(dy, a, out=base.clone(), accumulate=True)
        assert _rel(dw2, DY.T @ A + base.double().cpu()) < 2e-5


def test_tc_bias_not_16_byte_aligned(tc_ops):
    """The epilogue reads the bias with broadcast 16-byte loads when it can; a bias that is a view starting 4 bytes
    into a buffer (and an N whose last store box is clipped) must take the scalar loads and give the same result."""
    ops = tc_ops
    g = torch.Generator(device="cuda").manual_seed(5)
    M, N, K = 300, 328, 96
    a = torch.randn(M, K, generator=g, device="cuda").bfloat16()
    w = (torch.randn(N, K, generator=g, device="cuda") / K ** 0.5).bfloat16()
    big = torch.randn(N + 1, generator=g, device="cuda")
    bias = big[1:]
    assert bias.data_ptr() % 16 == 4
    want = a.double().cpu() @ w.double().cpu().T + bias.double().cpu()
    assert _rel(ops.gemm_fwd(a, w, bias), want) < 1e-2
    assert _rel(ops.gemm_fwd(a, w, bias, out_dtype=torch.float32), want) < 2e-5
    assert torch.equal(ops.gemm_fwd(a, w, bias), ops.gemm_fwd(a, w, bias.clone()))


def test_tc_column_block_views(tc_ops):
    """Operands/outputs that are column blocks of wider buffers (ld != width), as the layer
    code passes them (gate blocks of G, rows H.. of W_fused)."""
    ops = tc_ops
    g = torch.Generator(device="cuda").manual_seed(1)
    M, H = 640, 256
    G = torch.randn(M, 5 * H, generator=g, device="cuda").bfloat16()
    Wf = (torch.randn(6 * H, H, generator=g, device="cuda") / 16).bfloat16()
    u = torch.randn(M, H, generator=g, device="cuda").bfloat16()
    du = ops.gemm_dgrad(G, Wf[H:], out_dtype=torch.float32)
    assert _rel(du, G.double().cpu() @ Wf[H:].double().cpu()) < 2e-5
    blk = G[:, H:2 * H]
    d1 = ops.gemm_dgrad(blk, Wf[:H], out_dtype=torch.float32)
    assert _rel(d1, blk.double().cpu() @ Wf[:H].double().cpu()) < 2e-5
    dW = torch.zeros(6 * H, H, device="cuda")
    ops.gemm_wgrad(G, u, out=dW[H:])
    assert (dW[:H] == 0).all() and _rel(dW[H:], G.double().cpu().T @ u.double().cpu()) < 2e-5
    out = torch.zeros(M, 3 * H, device="cuda", dtype=torch.bfloat16)
    ops.gemm_fwd(u, Wf[:H], None, out=out[:, H:2 * H])
    assert (out[:, :H] == 0).all() and (out[:, 2 * H:] == 0).all()
    assert _rel(out[:, H:2 * H], u.double().cpu() @ Wf[:H].double().cpu().T) < 1e-2


def test_tc_linearity_at_full_size(tc_ops):
    """cfg2-sized projection (M=192000, K=1024, N=5120): no CPU oracle at this size, so use
    linearity: (A1+A2)W == A1 W + A2 W with operands exactly representable in bf16."""
    ops = tc_ops
    g = torch.Generator(device="cuda").manual_seed(2)
    M, K, N = 192000, 1024, 5120
    a1 = torch.randint(-4, 5, (M, K), generator=g, device="cuda").bfloat16()
    a2 = torch.randint(-4, 5, (M, K), generator=g, device="cuda").bfloat16()
    w = torch.randint(-2, 3, (N, K), generator=g, device="cuda").bfloat16()
    y1 = ops.gemm_fwd(a1, w, None, out_dtype=torch.float32)
    y2 = ops.gemm_fwd(a2, w, None, out_dtype=torch.float32)
    y12 = ops.gemm_fwd(a1 + a2, w, None, out_dtype=torch.float32)
    assert torch.equal(y12, y1 + y2)                              # small integers: exact in fp32
    rows = torch.randint(0, M, (64,), generator=g, device="cuda")
    ref = a1[rows].double().cpu() @ w.double().cpu().T
    assert torch.equal(y1[rows].double().cpu(), ref)
