"""torch.ops.graft.*: the TORCH_LIBRARY extension over libgraft.so's C ABI (recommendersystem_b200/torch_ext).
CPU part: the extension builds, loads, registers every operator with a Meta kernel (shape inference without a GPU), and
refuses CPU tensors.  GPU part (marked): values equal the ctypes path, autograd formulas match PyTorch, and
torch.compile(fullgraph=True) traces through the operators."""
import numpy as np
import pytest
import torch

from recommendersystem_b200 import torch_ext

OPS = ("jagged_offsets", "token_segments", "embedding_gather", "scatter_add_sorted", "gemm_bf16", "rmsnorm_fwd", "rmsnorm_bwd", "rmsnorm",
       "attn_fwd", "attn_bwd", "attention", "logits_ce_fwd")


@pytest.fixture(scope="module")
def ops():
    torch_ext.build()
    return torch_ext.load()


def test_operators_registered_with_meta_kernels(ops):
    for name in OPS:
        assert hasattr(ops, name), name
    m = lambda *shape, dtype=torch.float32: torch.empty(*shape, dtype=dtype, device="meta")  # noqa: E731
    y, rstd = ops.rmsnorm_fwd(m(10, 64), m(64), 1e-5)
    assert y.shape == (10, 64) and rstd.shape == (10,)
    assert ops.gemm_bf16(m(8, 16, dtype=torch.bfloat16), m(4, 16, dtype=torch.bfloat16), True, True, False).shape == (8, 4)
    assert ops.gemm_bf16(m(16, 8, dtype=torch.bfloat16), m(16, 4, dtype=torch.bfloat16), False, False, True).dtype == torch.float32
    assert ops.embedding_gather(m(100, 32), m(7, dtype=torch.int32)).shape == (7, 32)
    assert ops.scatter_add_sorted(m(7, 32), m(7, dtype=torch.int32), 100).shape == (100, 32)
    qkv = m(256, 4 * 64, dtype=torch.bfloat16)
    i32 = lambda n: m(n, dtype=torch.int32)  # noqa: E731
    out, lse = ops.attn_fwd(qkv, 2, 1, 2, i32(256), i32(256), i32(256))
    assert out.shape == (256, 128) and lse.shape == (2, 256)
    s, e = ops.token_segments(m(2, 64, dtype=torch.int32))
    assert s.shape == (256,) and e.shape == (256,)
    lse, tgt = ops.logits_ce_fwd(m(5, 64, dtype=torch.bfloat16), m(1000, 64, dtype=torch.bfloat16), i32(5))
    assert lse.shape == (5,) and tgt.shape == (5,)


def test_cpu_tensors_are_refused(ops):
    with pytest.raises((RuntimeError, NotImplementedError)):
        ops.rmsnorm_fwd(torch.zeros(4, 64), torch.ones(64), 1e-5)


@pytest.mark.gpu
def test_ops_match_ctypes_path_and_torch_autograd(ops):
    from recommendersystem_b200 import lib as L
    dev = "cuda:0"
    g = torch.Generator().manual_seed(0)
    # rmsnorm: values and gradients vs plain PyTorch
    x = torch.randn(300, 256, generator=g).to(dev).requires_grad_(True)
    sc = (1 + 0.1 * torch.randn(256, generator=g)).to(dev).requires_grad_(True)
    y = ops.rmsnorm(x, sc, 1e-5)
    ref = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-5) * sc
    assert float((y - ref).abs().max()) < 1e-5
    dy = torch.randn(300, 256, generator=g).to(dev)
    gx, gs = torch.autograd.grad(y, (x, sc), dy)
    rx, rs = torch.autograd.grad(ref, (x, sc), dy)
    assert float((gx - rx).norm() / rx.norm()) < 1e-5 and float((gs - rs).norm() / rs.norm()) < 1e-5
    # embedding gather (bit exact) and its deterministic scatter-add backward
    table = torch.randn(500, 128, generator=g).to(dev).requires_grad_(True)
    idx = torch.randint(-1, 499, (2000,), generator=g).int().to(dev)
    rows = ops.embedding_gather(table, idx)
    want = table[torch.where(idx == -1, 499, idx).long()]
    assert torch.equal(rows, want.detach())
    go = torch.randn(2000, 128, generator=g).to(dev)
    (gt,) = torch.autograd.grad(rows, table, go)
    (rt,) = torch.autograd.grad(want, table, go)
    assert float((gt - rt).norm() / rt.norm()) < 2e-6
    # GEMM == the ctypes path
    a = torch.randn(384, 256, generator=g).bfloat16().to(dev); b = torch.randn(200, 256, generator=g).bfloat16().to(dev)
    c = torch.empty(384, 200, device=dev, dtype=torch.bfloat16)
    L.gemm(a, b, c, 384, 200, 256)
    assert torch.equal(ops.gemm_bf16(a, b, True, True, False), c)
    # attention forward / backward == the ctypes path
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_gpu_kernels import make_attn_case
    rows_, row_len, H, G = 2, 512, 4, 2
    uid, tm, qkv = make_attn_case(rows_, row_len, H, G, seed=3)
    T = rows_ * row_len
    tm_tok = torch.empty(T, dtype=torch.int32, device=dev); L.interleave_i32(tm.reshape(-1), tm_tok)
    s1, e1 = ops.token_segments(uid)
    s2, e2 = L.token_segments(uid)
    assert torch.equal(s1, s2) and torch.equal(e1, e2)
    at = L.attn_args(qkv, H, G, rows_, row_len, tm_tok, s2, e2)
    out = torch.empty(T, H * 64, device=dev, dtype=torch.bfloat16); lse = torch.empty(H, T, device=dev)
    L.attn_fwd(at, out, lse)
    q = qkv.clone().requires_grad_(True)
    o2 = ops.attention(q, H, G, rows_, tm_tok, s1, e1)
    assert torch.equal(o2, out)
    do = torch.randn(T, H * 64, generator=g).bfloat16().to(dev)
    dq = torch.empty_like(qkv); delta = torch.empty(H, T, device=dev)
    L.attn_bwd(at, out, do, lse, dq, delta)
    (gq,) = torch.autograd.grad(o2, q, do)
    assert torch.equal(gq, dq)
    # fused logits + CE statistics
    h = torch.randn(64, 128, generator=g).bfloat16().to(dev); items = (torch.randn(1000, 128, generator=g) * 0.3).bfloat16().to(dev)
    tgt = torch.randint(0, 1000, (64,), generator=g).int().to(dev)
    lse_, tl = ops.logits_ce_fwd(h, items, tgt)
    logits = (h.float() @ items.float().t()).bfloat16().float()
    assert float((lse_ - torch.logsumexp(logits, 1)).abs().max()) < 1e-4
    assert torch.equal(tl, logits[torch.arange(64), tgt.long()])


@pytest.mark.gpu
def test_torch_compile_traces_through_the_operators(ops):
    """The reference wraps its model in torch.compile (transformer.py:677): custom operators with Meta kernels and autograd
    formulas compile with fullgraph=True (no graph break), forward and backward."""
    dev = "cuda:0"

    def block(x, scale, w):
        y = torch.ops.graft.rmsnorm(x, scale, 1e-5)
        return torch.ops.graft.gemm_bf16(y.to(torch.bfloat16), w, True, True, True).sum()

    g = torch.Generator().manual_seed(1)
    x = torch.randn(256, 128, generator=g).to(dev).requires_grad_(True)
    sc = torch.ones(128, device=dev, requires_grad=True)
    w = torch.randn(64, 128, generator=g).bfloat16().to(dev)
    eager = block(x, sc, w)
    compiled = torch.compile(block, fullgraph=True)(x, sc, w)
    assert float((eager - compiled).abs()) <= 1e-3 * abs(float(eager)) + 1e-3
