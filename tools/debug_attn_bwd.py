"""Isolates the backward mismatch of test_attention_large_scores: error of dq / dk / dv per head for several score scales,
with and without mask ids, against fp32 autograd."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_kernels as T
from recommendersystem_b200 import lib as L
DEV = "cuda:0"
rows, row_len, H, G = 2, 1024, 4, 2
for long_runs in (True, False):
    for zero_tm in (False, True):
        for sc in (1.0, 1.8, 3.0):
            uid, tm, qkv = T.make_attn_case(rows, row_len, H, G, seed=21, long_runs=long_runs)
            if zero_tm:
                tm = torch.zeros_like(tm)
            qkv = qkv.float(); qkv[:, :(H + G) * 64] *= sc; qkv = qkv.bfloat16()
            a, out, lse2 = T._attn_run(qkv, uid, tm, rows, row_len, H, G)
            q = qkv.clone().float().requires_grad_(True)
            ref_o, ref_lse = T.attn_reference(q, uid, tm, rows, row_len, H, G)
            do = T.g_randn(rows * row_len, H * 64, seed=11).bfloat16()
            ref_o.backward(do.float())
            dqkv = torch.empty_like(qkv); delta = torch.empty(H, rows * row_len, device=DEV)
            L.attn_bwd(a, out, do, lse2, dqkv, delta)
            g = q.grad
            e = lambda x, y: float((x.float() - y).norm() / y.norm().clamp_min(1e-20))
            dq = [round(e(dqkv[:, h * 64:(h + 1) * 64], g[:, h * 64:(h + 1) * 64]), 4) for h in range(H)]
            dk = [round(e(dqkv[:, (H + i) * 64:(H + i + 1) * 64], g[:, (H + i) * 64:(H + i + 1) * 64]), 4) for i in range(G)]
            dv = [round(e(dqkv[:, (H + G + i) * 64:(H + G + i + 1) * 64], g[:, (H + G + i) * 64:(H + G + i + 1) * 64]), 4) for i in range(G)]
            # the same gradient from a reference that rounds O to bf16 before delta (what any flash backward sees)
            o16 = ref_o.detach().bfloat16().float()
            print(f"long_runs={long_runs} zero_tm={zero_tm} scale={sc}: out {e(out, ref_o.detach()):.4f} dq {dq} dk {dk} dv {dv} |dq_ref| {float(g[:, :H*64].norm()):.3e} |dq| {float(dqkv[:, :H*64].float().norm()):.3e}", flush=True)
