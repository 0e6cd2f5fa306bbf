"""fp32 mode of the cloze step (config["precision"] = "fp32"): the reference step WITHOUT autocast, parity 1e-4.

Same autograd-node shape as model._ClozeStep (one Function: forward -> 4 losses, backward -> every parameter gradient) over
the fp32 entry points of libgraft (csrc/fp32_ref.cu: CUDA-core GEMM, attention, RoPE, SwiGLU, GELU, cross-entropy) and the
kernels the bf16 path shares (gather, scatter-add, RMSNorm, action embedding, row selection, loss reductions).  Nothing is fused
and nothing is rounded to bf16: this is the verification path BASELINE configs[0] (C1, fp32) and the north star's 1e-4
tolerance ask for, sized for small models; pre-train mode, softmax attention, replicated head only.
Reference lines: /root/reference/notebooks/Training/transformer.model.py (`ref:`)."""
from __future__ import annotations

import torch

from . import lib as L

ALL_MEDIUMS = [0, 1]


def _f32(shape, dev):
    return torch.empty(shape, dtype=torch.float32, device=dev)


class ClozeStepF32(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, d, evaluate, *params):
        cfg, P = model.config, model._params
        assert not P.lora and not P.fused and not cfg.get("vocab_parallel") and cfg.get("attention", "softmax") == "softmax", \
            "fp32 mode covers the pre-train step with the reference's softmax attention and a replicated head"
        W = dict(zip(P.names, params))
        dev = W["ae.w"].device
        B, S = d["userid"].shape
        N, T = B * S, 2 * B * S
        dm, H, G, I, nl = cfg["embed_dim"], cfg["num_heads"], cfg["num_kv_heads"], cfg["intermediate_dim"], cfg["num_layers"]
        assert dm // H == 64
        QKV = (H + 2 * G) * 64
        V0 = cfg["vocab_sizes"]["0_matchedid"]
        Vt = V0 + cfg["vocab_sizes"]["1_matchedid"]
        M = cfg["metadata_emb_size"]
        L.check_rows(d["userid"], count=False)
        sv = {}
        # item table = id rows + metadata . Wp^T + bp   (ref:139-145 / 165-169, all fp32)
        meta = model.item_embedding.metadata_embedding.embedding.weight.detach()
        table = W["ie.id"].detach().clone()
        L.gemm_f32(meta, W["ie.pw"], table, Vt + 1, dm, M, accumulate=True, bias=W["ie.pb"])
        # tokens (ref:464-467, 403-415)
        x = _f32((T, dm), dev)
        mid = d["matchedid"].reshape(-1)
        L.item_gather(table, mid, x)
        aa = model.action_embedding.kernel_args(d, W, N)
        aa.round_bf16 = 0
        feats = _f32((N, 32), dev)
        L.action_embed(aa, x, feats)
        sv.update(feats=feats, aa=aa, mid=mid, batch=d, table=table)
        tm_tok = torch.empty(T, dtype=torch.int32, device=dev); L.interleave_i32(d["token_mask_ids"].reshape(-1), tm_tok)
        seg_s, seg_e = L.token_segments(d["userid"])
        sv["seg"] = (tm_tok, seg_s, seg_e)
        cos, sin = model.transformers.freqs_cos, model.transformers.freqs_sin
        assert 2 * S <= cos.shape[0]
        layers = []
        for i in range(nl):   # ref:297-343
            xn = _f32((T, dm), dev); r1 = _f32((T,), dev)
            L.rmsnorm_fwd(x, W[f"l{i}.san"], None, xn, r1)
            qkv = _f32((T, QKV), dev)
            L.gemm_f32(xn, W[f"l{i}.q"], qkv, T, H * 64, dm, ldc=QKV)
            L.gemm_f32(xn, W[f"l{i}.k"], qkv[:, H * 64:], T, G * 64, dm, ldc=QKV)
            L.gemm_f32(xn, W[f"l{i}.v"], qkv[:, (H + G) * 64:], T, G * 64, dm, ldc=QKV)
            L.rope_f32(qkv, H + G, 2 * S, cos, sin)
            ao = _f32((T, H * 64), dev); lse = _f32((H, T), dev)
            L.attn_f32_fwd(qkv, H, G, B, 2 * S, tm_tok, seg_s, seg_e, ao, lse)
            h = x.clone()
            L.gemm_f32(ao, W[f"l{i}.o"], h, T, dm, H * 64, accumulate=True)          # h = x + attn(...)
            hn = _f32((T, dm), dev); r2 = _f32((T,), dev)
            L.rmsnorm_fwd(h, W[f"l{i}.mln"], None, hn, r2)
            gu = _f32((T, 2 * I), dev)
            L.gemm_f32(hn, W[f"l{i}.w1"], gu, T, I, dm, ldc=2 * I)
            L.gemm_f32(hn, W[f"l{i}.w3"], gu[:, I:], T, I, dm, ldc=2 * I)
            act = _f32((T, I), dev)
            L.swiglu_f32_fwd(gu, act)
            xo = h.clone()
            L.gemm_f32(act, W[f"l{i}.w2"], xo, T, dm, I, accumulate=True)            # x = h + mlp(h)
            layers.append((x, xn, r1, qkv, ao, lse, h, hn, r2, gu, act))
            x = xo
        hidden = _f32((T, dm), dev); rF = _f32((T,), dev)
        L.rmsnorm_fwd(x, W["norm"], None, hidden, rF)
        sv.update(layers=layers, xL=x, rF=rF)
        # heads (ref:499-529)
        cap = cfg["mask_topk"] * B
        out = torch.zeros(16, dtype=torch.float32, device=dev)
        heads = []
        for m in ALL_MEDIUMS:
            lo, hi = (0, V0) if m == 0 else (V0, Vt)
            K_m = hi - lo
            sel, ws, lab, pos, cnt = L.select_rows(d[f"{m}.watch.weight"].reshape(-1), d[f"{m}.watch.label"].reshape(-1),
                                                   d[f"{m}.watch.position"].reshape(-1).to(torch.int32), cap)
            hs = _f32((cap, dm), dev); L.gather_rows_f32(hidden, sel, 0, hs)
            logits = torch.zeros((cap, K_m), dtype=torch.float32, device=dev)
            L.gemm_f32(hs, table[lo:hi], logits, cap, K_m, dm, m_count=cnt)
            pmax = _f32((1, cap), dev); psum = _f32((1, cap), dev); tgt = _f32((cap,), dev)
            L.ce_rows_f32(logits, K_m, pos, pmax, psum, tgt)
            lse_r = _f32((cap,), dev); coef = _f32((cap,), dev); term = _f32((cap,), dev)
            L.ce_finish(pmax, psum, 1, tgt, lab, ws, lse_r, coef, term, out[2 * m:2 * m + 1])
            rsel, rws, rlab, _rp, rcnt = L.select_rows(d[f"{m}.rating.weight"].reshape(-1), d[f"{m}.rating.label"].reshape(-1),
                                                       d[f"{m}.rating.position"].reshape(-1).to(torch.int32), cap)
            rhs = _f32((cap, dm), dev); L.gather_rows_f32(hidden, rsel, 1, rhs)
            pre = _f32((cap, dm), dev)
            L.gemm_f32(rhs, W["rh.w0"], pre, cap, dm, dm, bias=W["rh.b0"])
            h1 = _f32((cap, dm), dev); L.gelu_f32(pre, h1)
            pred = _f32((cap,), dev)
            L.gemm_f32(h1, W["rh.w2"], pred, cap, 1, dm, bias=W["rh.b2"])
            loss4 = out[4 + 4 * m: 8 + 4 * m]
            L.rating_moments_f32(pred, rlab, rws, cfg["rating_mean"], loss4)
            heads.append(dict(lo=lo, hi=hi, sel=sel, hs=hs, pos=pos, lse=lse_r, coef=coef, cnt=cnt, logits=logits, rsel=rsel, rhs=rhs,
                              pre=pre, h1=h1, pred=pred, rlab=rlab, rws=rws))
        out[1] = out[4]
        out[3] = out[8]
        sv.update(heads=heads, out=out, dims=(B, S, N, T, dm, H, G, I, nl, QKV, V0, Vt, M, cap))
        ctx.sv, ctx.model = sv, model
        ctx.save_for_backward(*params)
        ctx.set_materialize_grads(False)
        res = out[:4].clone()
        if evaluate:
            mom = torch.stack((out[4:7], out[8:11])).clone()
            ctx.mark_non_differentiable(mom)
            return res, mom
        return res

    @staticmethod
    def backward(ctx, gout, *unused):
        sv, model = ctx.sv, ctx.model
        cfg, P = model.config, model._params
        W = dict(zip(P.names, ctx.saved_tensors))
        B, S, N, T, dm, H, G, I, nl, QKV, V0, Vt, M, cap = sv["dims"]
        dev = gout.device
        gout = gout.contiguous().float()
        G_ = {}
        table = sv["table"]
        dtable = torch.zeros((Vt + 1, dm), dtype=torch.float32, device=dev)
        dhid = torch.zeros((T, dm), dtype=torch.float32, device=dev)
        dw0 = torch.zeros((dm, dm), dtype=torch.float32, device=dev); db0 = torch.zeros(dm, dtype=torch.float32, device=dev)
        dw2 = torch.zeros((1, dm), dtype=torch.float32, device=dev); db2 = torch.zeros(1, dtype=torch.float32, device=dev)
        tmp_b0 = _f32((dm,), dev)
        for m, hd in enumerate(sv["heads"]):
            lo, hi = hd["lo"], hd["hi"]
            K_m = hi - lo
            coef = hd["coef"].clone(); L.scale_by_scalar(coef, gout[2 * m])
            dlog = hd["logits"]                                     # turned into coef (softmax - onehot) in place
            L.ce_grad_f32(dlog, K_m, hd["pos"], hd["lse"], coef)    # rows beyond the count: logits 0, coef 0 -> gradient 0
            dhs = _f32((cap, dm), dev)
            L.gemm_f32(dlog, table[lo:hi], dhs, cap, dm, K_m, b_nk=False)                       # d hidden rows = dlogits . items
            L.gemm_f32(dlog, hd["hs"], dtable[lo:hi], K_m, dm, cap, a_t=True, lda=K_m, b_nk=False)  # d items = dlogits^T . hidden rows
            L.scatter_rows_f32(dhs, hd["sel"], 0, dhid)
            # rating head (ref:355-359, 391-401)
            dp = _f32((cap,), dev)
            loss4 = sv["out"][4 + 4 * m: 8 + 4 * m]
            L.rating_dpred_f32(hd["pred"], hd["rlab"], hd["rws"], cfg["rating_mean"], loss4, gout[2 * m + 1:2 * m + 2], dp)
            L.gemm_f32(dp, hd["h1"], dw2, 1, dm, cap, a_t=True, lda=1, b_nk=False, accumulate=True)   # dW2 [1, d] += dp^T . h1
            db2 += dp.sum()
            dh1 = _f32((cap, dm), dev)
            L.gemm_f32(dp, W["rh.w2"], dh1, cap, dm, 1, lda=1, b_nk=False, ldb=dm)                # dh1 = dp (x) w2
            dpre = _f32((cap, dm), dev); L.gelu_f32(hd["pre"], dpre, dy=dh1)
            L.gemm_f32(dpre, hd["rhs"], dw0, dm, dm, cap, a_t=True, lda=dm, b_nk=False, accumulate=True)
            L.colsum(dpre, tmp_b0); db0 += tmp_b0
            drhs = _f32((cap, dm), dev)
            L.gemm_f32(dpre, W["rh.w0"], drhs, cap, dm, dm, b_nk=False)
            L.scatter_rows_f32(drhs, hd["rsel"], 1, dhid)
        G_["rh.w0"], G_["rh.b0"], G_["rh.w2"], G_["rh.b2"] = dw0, db0, dw2, db2
        dx = _f32((T, dm), dev); dn = _f32((dm,), dev)
        L.rmsnorm_bwd(sv["xL"], W["norm"], sv["rF"], None, dhid, None, dx, None, dn)
        G_["norm"] = dn
        tm_tok, seg_s, seg_e = sv["seg"]
        cos, sin = model.transformers.freqs_cos, model.transformers.freqs_sin
        for i in reversed(range(nl)):
            x, xn, r1, qkv, ao, lse, h, hn, r2, gu, act = sv["layers"][i]
            dact = _f32((T, I), dev)
            L.gemm_f32(dx, W[f"l{i}.w2"], dact, T, I, dm, b_nk=False)
            g_w2 = _f32((dm, I), dev)
            L.gemm_f32(dx, act, g_w2, dm, I, T, a_t=True, lda=dm, b_nk=False)
            dgu = _f32((T, 2 * I), dev)
            L.swiglu_f32_bwd(gu, dact, dgu)
            dhn = _f32((T, dm), dev)
            L.gemm_f32(dgu, W[f"l{i}.w1"], dhn, T, dm, I, lda=2 * I, b_nk=False)
            L.gemm_f32(dgu[:, I:], W[f"l{i}.w3"], dhn, T, dm, I, lda=2 * I, b_nk=False, accumulate=True)
            g_w1 = _f32((I, dm), dev); g_w3 = _f32((I, dm), dev)
            L.gemm_f32(dgu, hn, g_w1, I, dm, T, a_t=True, lda=2 * I, b_nk=False)
            L.gemm_f32(dgu[:, I:], hn, g_w3, I, dm, T, a_t=True, lda=2 * I, b_nk=False)
            dh = _f32((T, dm), dev); g_mln = _f32((dm,), dev)
            L.rmsnorm_bwd(h, W[f"l{i}.mln"], r2, None, dhn, dx, dh, None, g_mln)          # dh = dx + norm_bwd(dhn)
            dao = _f32((T, H * 64), dev)
            L.gemm_f32(dh, W[f"l{i}.o"], dao, T, H * 64, dm, b_nk=False)
            g_o = _f32((dm, H * 64), dev)
            L.gemm_f32(dh, ao, g_o, dm, H * 64, T, a_t=True, lda=dm, b_nk=False)
            dqkv = _f32((T, QKV), dev); delta = _f32((H, T), dev)
            L.attn_f32_bwd(qkv, H, G, B, 2 * S, tm_tok, seg_s, seg_e, ao, dao, lse, dqkv, delta)
            L.rope_f32(dqkv, H + G, 2 * S, cos, sin, inverse=True)
            dxn = _f32((T, dm), dev)
            L.gemm_f32(dqkv, W[f"l{i}.q"], dxn, T, dm, H * 64, lda=QKV, b_nk=False)
            L.gemm_f32(dqkv[:, H * 64:], W[f"l{i}.k"], dxn, T, dm, G * 64, lda=QKV, b_nk=False, accumulate=True)
            L.gemm_f32(dqkv[:, (H + G) * 64:], W[f"l{i}.v"], dxn, T, dm, G * 64, lda=QKV, b_nk=False, accumulate=True)
            g_q = _f32((H * 64, dm), dev); g_k = _f32((G * 64, dm), dev); g_v = _f32((G * 64, dm), dev)
            L.gemm_f32(dqkv, xn, g_q, H * 64, dm, T, a_t=True, lda=QKV, b_nk=False)
            L.gemm_f32(dqkv[:, H * 64:], xn, g_k, G * 64, dm, T, a_t=True, lda=QKV, b_nk=False)
            L.gemm_f32(dqkv[:, (H + G) * 64:], xn, g_v, G * 64, dm, T, a_t=True, lda=QKV, b_nk=False)
            g_san = _f32((dm,), dev)
            dx = _f32((T, dm), dev)
            L.rmsnorm_bwd(x, W[f"l{i}.san"], r1, None, dxn, dh, dx, None, g_san)          # dx = dh + norm_bwd(dxn)
            G_.update({f"l{i}.q": g_q, f"l{i}.k": g_k, f"l{i}.v": g_v, f"l{i}.o": g_o, f"l{i}.w1": g_w1, f"l{i}.w3": g_w3, f"l{i}.w2": g_w2,
                       f"l{i}.san": g_san, f"l{i}.mln": g_mln})
            sv["layers"][i] = None
        # token embeddings
        L.scatter_add_sorted(dx, 2 * dm, sv["mid"], dm, dtable)
        ae = {k: torch.zeros_like(W[k]) for k in ("ae.w", "ae.b", "ae.gender", "ae.source", "ae.status", "ae.cos", "ae.sin")}
        L.action_embed_bwd(sv["aa"], sv["feats"], dx, ae["ae.w"], ae["ae.b"], ae["ae.gender"], ae["ae.source"], ae["ae.status"],
                           ae["ae.cos"], ae["ae.sin"])
        G_.update(ae)
        G_["ie.id"] = dtable
        dpb = _f32((dm,), dev); L.colsum(dtable, dpb)
        G_["ie.pb"] = dpb
        meta = model.item_embedding.metadata_embedding.embedding.weight.detach()
        dpw = _f32((dm, M), dev)
        L.gemm_f32(dtable, meta, dpw, dm, M, Vt + 1, a_t=True, lda=dm, b_nk=False)     # dWp = dtable^T . metadata
        G_["ie.pw"] = dpw
        ctx.sv = None
        return (None, None, None, *[(G_[n].contiguous() if need else None) for n, need in zip(P.names, ctx.needs_input_grad[3:])])
