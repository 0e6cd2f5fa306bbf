import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from recommendersystem_b200 import lib as L
DEV = "cuda:0"
g = torch.Generator().manual_seed(0)
for M, N, K in ((128, 75, 128), (128, 150, 128), (200, 333, 128), (128, 100, 128)):
    h = (torch.randn(M, K, generator=g)).bfloat16().to(DEV)
    items_full = (torch.randn(N + 50, K, generator=g) * 0.3).bfloat16().to(DEV)
    items = items_full[25:25 + N]
    tgt_idx = torch.randint(-20, N + 20, (M,), generator=g).int().to(DEV)
    logits = (h.float() @ items.float().t()).bfloat16().float()
    nb = L.ce_nblocks(N)
    pmax = torch.full((nb, M), 7.0, device=DEV); psum = torch.full((nb, M), 7.0, device=DEV); tgt = torch.zeros(M, device=DEV)
    L.gemm(h, items, None, M, N, K, epilogue=L.EPI_CE_PARTIAL, ce=dict(pmax=pmax, psum=psum, tgt=tgt, target=tgt_idx))
    mo = torch.empty(M, device=DEV); so = torch.empty(M, device=DEV)
    L.ce_merge_partials(pmax, psum, nb, mo, so)
    lse = mo + so.log()
    ref = torch.logsumexp(logits, 1)
    print(M, N, "lse err", float((lse - ref).abs().max()), "pmax", pmax[:, 0].tolist(), "psum", psum[:, 0].tolist(), "finite", bool(torch.isfinite(lse).all()))
    inr = (tgt_idx >= 0) & (tgt_idx < N)
    want = torch.where(inr, logits[torch.arange(M), tgt_idx.clamp(0, N - 1).long()], torch.zeros_like(tgt))
    print("   tgt err", float((tgt - want).abs().max()))
