"""Drop-in `RecommenderModel` whose train step runs on libgraft's sm_100a kernels.

Mirrors the call surface of the reference module
(/root/reference/notebooks/Training/transformer.model.py, cited as `ref:` below):
  * names: ALL_MEDIUMS, ALL_METRICS, RecommenderModel(config)          ref:1-2,346
  * model(d, evaluate) -> [0.watch, 0.rating, 1.watch, 1.rating]        ref:493-529
    (rating entries are lists of 3 moments when evaluate=True);
    `d` is reshaped and masked IN PLACE exactly as ref:495-497 does, because the
    trainer reads d[...weight].sum() afterwards (transformer.py:261)
  * identical parameter tree / state_dict keys / init                    ref:5-12,346-360
  * load_pretrained_embeddings semantics                                 ref:379-389

Host side is PyTorch for parameters, autograd plumbing and allocation only.
Every contraction, gather, scatter, norm and loss of the step is a libgraft
kernel reached through recommendersystem_b200.lib (ctypes over the C ABI).
There is no fallback path: without CUDA + libgraft.so the forward raises.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import comm
from . import lib as L

ALL_MEDIUMS = [0, 1]
ALL_METRICS = ["watch", "rating"]


def init_weights(module, std=0.006):
    """ref:5-12."""
    if isinstance(module, nn.Linear):
        module.weight.data.normal_(mean=0.0, std=std)
        if module.bias is not None:
            module.bias.data.zero_()
    elif isinstance(module, nn.Embedding):
        module.weight.data.normal_(mean=0.0, std=std)
        module.weight.data[-1].fill_(0)


def _need_cuda(t, what):
    if not t.is_cuda:
        raise L.GraftError(f"{what} has no CPU path: move the module and its inputs to a CUDA device")


_BATCH_INT = ("userid", "token_mask_ids", "gender", "source", "matchedid", "status", "rope_input_pos")


def _normalise_batch(d, device=None):
    """Element types the kernels hard-code (ids Int32, time Float64, the rest Float32; `.position` is
    int64 in the finetune loader, transformer.py:182-183).  The reference module indexes with whatever integer dtype it is
    given; here a mismatch would be reinterpreted bit-wise, so every entry is converted (a no-op when it already
    matches) and the dict is updated IN PLACE, like the step's other edits of `d`."""
    for k in list(d):
        v = d[k]
        if k == "time":
            want = torch.float64
        elif k in _BATCH_INT or k.endswith(".position"):
            want = torch.int32
        else:
            want = torch.float32
        if v.dtype != want or (device is not None and v.device != device):
            d[k] = v.to(device=device if device is not None else v.device, dtype=want)
    return d


# ------------------------------------------------------------------------------
# Parameter containers.  They carry the reference's module/parameter names so
# that checkpoints interchange.  The training step never calls
# their forward (it is one fused autograd node, _ClozeStep); the embedding
# modules still answer a direct call the way the reference's do, through the
# same libgraft kernels and without autograd (Finetune/register.py:29 exports
# the item table with `model.item_embedding(torch.arange(...))`).
# ------------------------------------------------------------------------------
class MaskedEmbedding(nn.Module):
    def __init__(self, vocab_size, embed_dim):
        super().__init__()
        self.vocab_size = vocab_size + 1  # last row = the "-1" (masked) row, ref:15-24
        self.embedding = nn.Embedding(self.vocab_size, embed_dim)

    @torch.no_grad()
    def forward(self, x):
        """ref:23-24: rows of the table, -1 -> the last row.  Inference-only (no autograd)."""
        w = self.embedding.weight
        _need_cuda(w, "MaskedEmbedding.forward")
        return L.embedding_gather(w.detach(), x.to(w.device))


class ActionEmbedding(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.periodic_time_cos = nn.Parameter(torch.zeros(2))
        self.periodic_time_sin = nn.Parameter(torch.zeros(2))
        vs = config["vocab_sizes"]
        self.status_embedding = MaskedEmbedding(vs["status"], 16)
        self.gender_embedding = MaskedEmbedding(vs["gender"], 4)
        self.source_embedding = MaskedEmbedding(vs["source"], 4)
        self.linear = nn.Linear(32, config["embed_dim"])  # ref:36-49: 1+2+2+4+4+1+1+16+1 features
        self.config = config

    def kernel_args(self, d, W, N):
        """graft_action_args for batch `d` (element types checked: the kernels hard-code them)."""
        cfg = self.config
        a = L.ActionArgs()
        a.time = L.tptr(d["time"], torch.float64, "time")
        a.gender, a.source, a.status = (L.tptr(d[k], torch.int32, k) for k in ("gender", "source", "status"))
        a.rating, a.progress = L.tptr(d["rating"], torch.float32, "rating"), L.tptr(d["progress"], torch.float32, "progress")
        a.periodic_cos, a.periodic_sin = L.ptr(W["ae.cos"]), L.ptr(W["ae.sin"])
        a.gender_table, a.source_table, a.status_table = L.ptr(W["ae.gender"]), L.ptr(W["ae.source"]), L.ptr(W["ae.status"])
        a.gender_rows, a.source_rows, a.status_rows = W["ae.gender"].shape[0], W["ae.source"].shape[0], W["ae.status"].shape[0]
        a.weight, a.bias = L.ptr(W["ae.w"]), L.ptr(W["ae.b"])
        a.min_ts, a.max_ts = float(cfg["min_ts"]), float(cfg["max_ts"])
        a.rating_mean, a.rating_std = float(cfg["rating_mean"]), float(cfg["rating_std"])
        a.n, a.d, a.round_bf16 = N, cfg["embed_dim"], 1
        return a

    @torch.no_grad()
    def forward(self, d):
        """ref:51-94 called on its own: [*, d] bf16 (what the autocast reference returns).  Inference-only."""
        w = self.linear.weight
        _need_cuda(w, "ActionEmbedding.forward")
        d = _normalise_batch({k: d[k] for k in ("time", "gender", "source", "status", "rating", "progress")}, w.device)
        shape = d["time"].shape
        N = d["time"].numel()
        W = {"ae.cos": self.periodic_time_cos, "ae.sin": self.periodic_time_sin, "ae.w": w, "ae.b": self.linear.bias,
             "ae.status": self.status_embedding.embedding.weight, "ae.gender": self.gender_embedding.embedding.weight,
             "ae.source": self.source_embedding.embedding.weight}
        W = {k: v.detach() for k, v in W.items()}
        x = torch.empty((2 * N, w.shape[0]), dtype=torch.float32, device=w.device)
        L.action_embed(self.kernel_args(d, W, N), x, torch.empty((N, 32), dtype=torch.float32, device=w.device))
        return x[1::2].to(torch.bfloat16).view(*shape, w.shape[0])


class ItemEmbedding(nn.Module):
    """ref:96-145.  Un-fused: id table + frozen metadata table + projection.  Fused
    (finetune inference from construction; finetune training after the checkpoint is loaded,
    ref:116-133): one [V+1, d] buffer `fused_embedding`, computed in fp32 exactly as
    ItemEmbedding.fuse() does - a load-time step, not part of the training step."""

    def __init__(self, config):
        super().__init__()
        vocab = sum(config["vocab_sizes"][f"{m}_matchedid"] for m in ALL_MEDIUMS)
        self.vocab_size = vocab
        self.fused = bool(config["finetune"]) and config["forward"] == "inference"
        if self.fused:
            self.register_buffer("fused_embedding", torch.empty(vocab + 1, config["embed_dim"]))
            return
        self.matchedid_embedding = MaskedEmbedding(vocab, config["embed_dim"])
        self.metadata_embedding = MaskedEmbedding(vocab, config["metadata_emb_size"])
        for p in self.metadata_embedding.parameters():
            p.requires_grad = False  # ref:113-114
        self.projection_layer = nn.Linear(config["metadata_emb_size"], config["embed_dim"])
        self.register_load_state_dict_pre_hook(self.reject_fused_checkpoint)
        if config["finetune"]:
            self.register_load_state_dict_post_hook(lambda module, _: module.fuse())

    def fuse(self):
        if self.fused:
            return
        with torch.no_grad():
            fused = self.matchedid_embedding.embedding.weight + torch.nn.functional.linear(
                self.metadata_embedding.embedding.weight, self.projection_layer.weight, self.projection_layer.bias)
        self.register_buffer("fused_embedding", fused)
        del self.matchedid_embedding
        del self.metadata_embedding
        del self.projection_layer
        self.fused = True

    @staticmethod
    def reject_fused_checkpoint(module, state_dict, prefix, *args):
        assert (prefix + "fused_embedding") not in state_dict

    # ---- kernels' view of the tables --------------------------------------------------------
    _weights_version = 0   # bumped by optim.FlatAdamW.step: its raw-pointer updates do not touch tensor._version
    _meta_cache = None
    _table_cache = None

    def meta_bf16(self):
        """bf16 copy of the frozen metadata table, rows zero-padded to a multiple of 64 columns: the row pitch is then a
        multiple of 128 bytes, so every 128-byte row of a TMA box is one aligned line (a 16-byte-aligned pitch made each
        box row straddle two lines: twice the requests for the 1.2 GB operand of the item-table GEMMs)."""
        w = self.metadata_embedding.embedding.weight
        key = (w.data_ptr(), w._version, str(w.device))
        if self._meta_cache is None or self._meta_cache[0] != key:
            Mp = (w.shape[1] + 63) // 64 * 64
            buf = torch.empty((w.shape[0], Mp), dtype=torch.bfloat16, device=w.device)
            L.cast_pad_rows(w.detach(), buf)
            self._meta_cache = (key, buf, Mp)
        return self._meta_cache[1], self._meta_cache[2]

    @torch.no_grad()
    def table(self):
        """[V+1, d] fp32 = ItemEmbedding.fuse() (ref:120-133): id rows + bf16(metadata . Wp^T + bp); the buffer itself when
        fused.  Cached until a weight changes (tensor versions + the optimiser's step counter)."""
        if self.fused:
            return self.fused_embedding
        ew, pw, pb = self.matchedid_embedding.embedding.weight, self.projection_layer.weight, self.projection_layer.bias
        _need_cuda(ew, "ItemEmbedding")
        key = (ew.data_ptr(), ew._version, pw.data_ptr(), pw._version, pb._version, ItemEmbedding._weights_version, str(ew.device))
        if self._table_cache is None or self._table_cache[0] != key:
            meta16, Mp = self.meta_bf16()
            dm = ew.shape[1]
            pw16 = torch.empty((dm, Mp), dtype=torch.bfloat16, device=ew.device); L.cast_pad_rows(pw.detach(), pw16)
            tab = torch.empty(ew.shape, dtype=torch.float32, device=ew.device)
            L.gemm(meta16, pw16, tab, ew.shape[0], dm, Mp, epilogue=L.EPI_TABLE, resid=ew.detach(), bias=pb.detach())
            self._table_cache = (key, tab)
        return self._table_cache[1]

    @torch.no_grad()
    def forward(self, x):
        """ref:139-145 called on its own (Finetune/register.py:29): rows of the fused table, [*x.shape, d] fp32.
        Inference-only: the training step differentiates the table inside _ClozeStep."""
        tab = self.table()
        _need_cuda(tab, "ItemEmbedding.forward")
        return L.embedding_gather(tab, x.to(tab.device))


class DualItemEmbedding(nn.Module):
    """Weight-tied cloze head: registers the same ItemEmbedding again (ref:148-151,354),
    which is why state_dict carries watch_head.item_embedding.* aliases."""

    def __init__(self, item_embedding, v0=None):
        super().__init__()
        self.item_embedding = item_embedding
        self._v0 = v0  # catalog size of medium 0 (manga): the tied table holds medium 0 rows first (ref:155-161)

    @torch.no_grad()
    def forward(self, inputs):
        """ref:153-170 called on its own: (x [rows, d], medium) -> logits [rows, K_medium] bf16 = x . items^T with the
        catalog slice of the tied item table (no bias, no temperature).  Inference-only; the training step never
        materialises this matrix (cross-entropy is fused into the logit GEMM's epilogue)."""
        x, medium = inputs
        ie = self.item_embedding
        tab = ie.table()
        _need_cuda(tab, "DualItemEmbedding.forward")
        V0 = self._v0
        assert medium in (0, 1)
        lo, hi = (0, V0) if medium == 0 else (V0, ie.vocab_size)
        x2 = x.reshape(-1, x.shape[-1]).to(tab.device)
        rows, dm, K = x2.shape[0], x2.shape[1], hi - lo
        x16 = x2.to(torch.bfloat16).contiguous()
        items16 = torch.empty((K, dm), dtype=torch.bfloat16, device=tab.device); L.cast_bf16(tab[lo:hi], items16)
        ldc = (K + 7) // 8 * 8
        out = torch.empty((rows, ldc), dtype=torch.bfloat16, device=tab.device)
        L.gemm(x16, items16, out, rows, K, dm, ldc=ldc)
        return out[:, :K].view(*x.shape[:-1], K)


class RMSNorm(nn.Module):
    def __init__(self, dim, eps=1e-5):
        super().__init__()
        self.eps = eps
        self.scale = nn.Parameter(torch.ones(dim))


class SwiGLU(nn.Module):
    def __init__(self, dim, inter):
        super().__init__()
        self.w1 = nn.Linear(dim, inter, bias=False)
        self.w2 = nn.Linear(inter, dim, bias=False)
        self.w3 = nn.Linear(dim, inter, bias=False)


class FlashAttention(nn.Module):  # name kept for state_dict/module-tree parity (ref:216)
    def __init__(self, config):
        super().__init__()
        d, H, G = config["embed_dim"], config["num_heads"], config["num_kv_heads"]
        hd = d // H
        self.q_proj = nn.Linear(d, H * hd, bias=False)
        self.k_proj = nn.Linear(d, G * hd, bias=False)
        self.v_proj = nn.Linear(d, G * hd, bias=False)
        self.output_proj = nn.Linear(H * hd, d, bias=False)
        if config["finetune"]:  # LoRA r=8, alpha=16 on q and v, ref:235-254
            self.q_proj_lora_A = nn.Linear(d, 8, bias=False)
            self.q_proj_lora_B = nn.Linear(8, H * hd, bias=False)
            self.v_proj_lora_A = nn.Linear(d, 8, bias=False)
            self.v_proj_lora_B = nn.Linear(8, G * hd, bias=False)
            # (the reference zeroes *_lora_B here, ref:252-254, but RecommenderModel's
            # self.apply(init_weights), ref:360, re-draws every Linear afterwards: B is N(0, .006))


class TransformerBlock(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.attn = FlashAttention(config)
        self.mlp = SwiGLU(config["embed_dim"], config["intermediate_dim"])
        self.sa_norm = RMSNorm(config["embed_dim"])
        self.mlp_norm = RMSNorm(config["embed_dim"])


class Transformer(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.layers = nn.ModuleList([TransformerBlock(config) for _ in range(config["num_layers"])])
        self.norm = RMSNorm(config["embed_dim"])
        hd = config["embed_dim"] // config["num_heads"]
        n = 2 * config["max_sequence_length"]  # item + action token per position, ref:321-322
        inv = 1.0 / (500000.0 ** (torch.arange(0, hd, 2)[: hd // 2].float() / hd))  # ref:173-179
        ang = torch.outer(torch.arange(n, dtype=torch.float32), inv)
        self.register_buffer("freqs_cos", torch.cos(ang), persistent=False)
        self.register_buffer("freqs_sin", torch.sin(ang), persistent=False)
        # the same tables transposed ([hd/2, n]) for the qkv GEMM's RoPE epilogue: a warp's 32 rows are consecutive
        # positions, so it reads one 128-byte line per (pair, table) instead of 32 scattered 16-byte pieces
        self.register_buffer("freqs_cos_t", torch.cos(ang).t().contiguous(), persistent=False)
        self.register_buffer("freqs_sin_t", torch.sin(ang).t().contiguous(), persistent=False)


# ------------------------------------------------------------------------------
class _Params:
    """Flat, ordered view of the trainable tensors the step differentiates."""

    def __init__(self, model):
        ae, ie, tr = model.action_embedding, model.item_embedding, model.transformers
        self.names, self.tensors = [], []

        def add(n, t):
            self.names.append(n)
            self.tensors.append(t)

        add("ae.cos", ae.periodic_time_cos); add("ae.sin", ae.periodic_time_sin)
        add("ae.status", ae.status_embedding.embedding.weight)
        add("ae.gender", ae.gender_embedding.embedding.weight)
        add("ae.source", ae.source_embedding.embedding.weight)
        add("ae.w", ae.linear.weight); add("ae.b", ae.linear.bias)
        self.fused = ie.fused
        self.lora = bool(model.config["finetune"])
        if not ie.fused:
            add("ie.id", ie.matchedid_embedding.embedding.weight)
            add("ie.pw", ie.projection_layer.weight); add("ie.pb", ie.projection_layer.bias)
        for i, l in enumerate(tr.layers):
            if self.lora:
                add(f"l{i}.qA", l.attn.q_proj_lora_A.weight); add(f"l{i}.qB", l.attn.q_proj_lora_B.weight)
                add(f"l{i}.vA", l.attn.v_proj_lora_A.weight); add(f"l{i}.vB", l.attn.v_proj_lora_B.weight)
            add(f"l{i}.q", l.attn.q_proj.weight); add(f"l{i}.k", l.attn.k_proj.weight)
            add(f"l{i}.v", l.attn.v_proj.weight); add(f"l{i}.o", l.attn.output_proj.weight)
            add(f"l{i}.w1", l.mlp.w1.weight); add(f"l{i}.w2", l.mlp.w2.weight); add(f"l{i}.w3", l.mlp.w3.weight)
            add(f"l{i}.san", l.sa_norm.scale); add(f"l{i}.mln", l.mlp_norm.scale)
        add("norm", tr.norm.scale)
        add("rh.w0", model.rating_head[0].weight); add("rh.b0", model.rating_head[0].bias)
        add("rh.w2", model.rating_head[2].weight); add("rh.b2", model.rating_head[2].bias)
        self.index = {n: i for i, n in enumerate(self.names)}


def _attn_mode(cfg):
    """config["attention"]: "softmax" (the reference, default) or "silu" (HSTU-style point-wise attention,
    BASELINE configs[2]; not in the reference - defined by oracle/cloze_oracle.py:silu_pointwise_attention).
    config["attention_safe"] (softmax only): forward with the online-softmax kernel - exact for any score range; the
    default fixed-reference kernel is exact while no score exceeds its row's first-tile maximum by > 69 nats and
    raises lib.FLAG_ATTN_RANGE otherwise (lib.check_status())."""
    kind = cfg.get("attention", "softmax")
    assert kind in ("softmax", "silu"), kind
    if kind == "silu":
        return L.ATTN_SILU
    return L.ATTN_SOFTMAX | (L.ATTN_SAFE if cfg.get("attention_safe") else 0)


def _bf16(shape, dev):
    return torch.empty(shape, dtype=torch.bfloat16, device=dev)


def _f32(shape, dev):
    return torch.empty(shape, dtype=torch.float32, device=dev)


def _block_weights_bf16(wq_, wk_, wv_, wo_, w1_, w3_, w2_, H, G, dm, I, dev, with_o_t=False):
    """bf16 operand copies of one block's weights in ONE launch (graft_cast_jobs): fused [q; k; v], Wo (+ its transpose for
    the backward), w1 / w3 interleaved in 64-row blocks (one GEMM tile then holds matching gate and up columns and the SwiGLU
    product is formed in its epilogue), w2."""
    QKV = (H + 2 * G) * 64
    qkv_w = _bf16((QKV, dm), dev); o_w = _bf16((dm, H * 64), dev); w13 = _bf16((2 * I, dm), dev); w2 = _bf16((dm, I), dev)
    o_t = _bf16((H * 64, dm), dev) if with_o_t else None
    det = lambda t: t.detach().contiguous()
    jobs = [(det(wq_), qkv_w, L.CAST_PLAIN, 0), (det(wk_), qkv_w, L.CAST_PLAIN, H * 64), (det(wv_), qkv_w, L.CAST_PLAIN, (H + G) * 64),
            (det(wo_), o_w, L.CAST_PLAIN, 0), (det(w1_), w13, L.CAST_GATE64, 0), (det(w3_), w13, L.CAST_UP64, 0), (det(w2_), w2, L.CAST_PLAIN, 0)]
    if with_o_t:
        jobs.append((det(wo_), o_t, L.CAST_TRANSPOSE, 0))
    L.cast_jobs(jobs)
    return qkv_w, o_w, w13, w2, o_t


class _ClozeStep(torch.autograd.Function):
    """Whole step as one autograd node: forward returns the loss vector, backward
    consumes d(loss vector) (a device tensor, no host sync) and returns every
    parameter gradient.  All compute is libgraft kernels on the current stream."""

    @staticmethod
    def forward(ctx, model, d, evaluate, *params):
        cfg = model.config
        P = model._params
        W = dict(zip(P.names, params))
        dev = W["ae.w"].device
        if dev.type != "cuda":
            raise L.GraftError("the cloze step has no CPU path: move the model and batch to a CUDA device")
        B, S = d["userid"].shape
        N, T = B * S, 2 * B * S
        dm, H, G, I = cfg["embed_dim"], cfg["num_heads"], cfg["num_kv_heads"], cfg["intermediate_dim"]
        nl = cfg["num_layers"]
        assert dm // H == 64, "libgraft attention is built for head_dim 64 (as the reference configs use)"
        assert I % 64 == 0, "intermediate_dim must be a multiple of 64 (fused gate/up tile layout)"
        assert (2 * S) % 128 == 0, "rows must hold a multiple of 64 positions (128-token attention tiles)"
        # the reference raises a broadcast error when a row is longer than the RoPE table (ref:182-190, 321-322)
        assert 2 * S <= model.transformers.freqs_cos.shape[0], "row longer than the RoPE table (2 * max_sequence_length)"
        # run-based visibility == the reference's userid equality only if no userid recurs inside a row:
        # checked on the device, reported through the status word (lib.check_status) - no host sync in the step
        L.check_rows(d["userid"], count=False)
        QKV = (H + 2 * G) * 64
        V0 = cfg["vocab_sizes"]["0_matchedid"]
        Vt = V0 + cfg["vocab_sizes"]["1_matchedid"]
        sv = {}  # saved for backward

        # ---- bf16 operand copies of the weights (autocast's casts, done once per step): one launch per block ----
        wq = []
        for i in range(nl):
            wq.append(_block_weights_bf16(W[f"l{i}.q"], W[f"l{i}.k"], W[f"l{i}.v"], W[f"l{i}.o"], W[f"l{i}.w1"], W[f"l{i}.w3"],
                                          W[f"l{i}.w2"], H, G, dm, I, dev, with_o_t=True))
        sv["wq"] = wq
        lw = []
        if P.lora:
            for i in range(nl):
                Aq = torch.zeros((64, dm), dtype=torch.bfloat16, device=dev); Aq[:8].copy_(W[f"l{i}.qA"])
                Av = torch.zeros((64, dm), dtype=torch.bfloat16, device=dev); Av[:8].copy_(W[f"l{i}.vA"])
                Bq2 = torch.zeros((H * 64, 64), dtype=torch.bfloat16, device=dev); Bq2[:, :8].copy_(2.0 * W[f"l{i}.qB"])
                Bv2 = torch.zeros((G * 64, 64), dtype=torch.bfloat16, device=dev); Bv2[:, :8].copy_(2.0 * W[f"l{i}.vB"])
                lw.append((Aq, Bq2, Av, Bv2))
        sv["lw"] = lw

        # ---- fused item table = id rows + bf16(meta . Wp^T + bp)   (ref:120-128 / 165-169) ----
        if P.fused:  # finetune: frozen, fused once at load time (ref:116-133)
            table, table16 = model._fused_tables()
            Mp = 0
        else:
            meta16, Mp = model.item_embedding.meta_bf16()
            pw16 = _bf16((dm, Mp), dev); L.cast_pad_rows(W["ie.pw"], pw16)
            table = _f32((Vt + 1, dm), dev)
            L.gemm(meta16, pw16, table, Vt + 1, dm, Mp, epilogue=L.EPI_TABLE, resid=W["ie.id"], bias=W["ie.pb"])
            table16 = _bf16((Vt + 1, dm), dev); L.cast_bf16(table, table16)
            sv["pw16"] = pw16
        sv["table16"] = table16

        # ---- tokens: x0[2i] = table[id], x0[2i+1] = action Linear (ref:464-467) ----
        x = _f32((T, dm), dev)
        mid = d["matchedid"].reshape(-1)
        L.item_gather(table, mid, x)
        aa = model.action_embedding.kernel_args(d, W, N)
        feats = _f32((N, 32), dev)
        L.action_embed(aa, x, feats)
        sv["feats"], sv["aa"], sv["mid"] = feats, aa, mid
        sv["batch"] = d  # `aa` holds raw device pointers into the batch: keep its tensors alive until backward

        tm_tok = torch.empty(T, dtype=torch.int32, device=dev); L.interleave_i32(d["token_mask_ids"].reshape(-1), tm_tok)
        seg_s, seg_e = L.token_segments(d["userid"])
        sv["seg"] = (tm_tok, seg_s, seg_e)
        cos, sin = model.transformers.freqs_cos, model.transformers.freqs_sin

        # ---- block stack (ref:297-343) ----
        layers = []
        dlt = None  # bf16 output of the previous block's w2 GEMM: its residual add rides on the next norm
        for i in range(nl):
            qkv_w, o_w, w13, w2, _o_t = wq[i]
            xn = _bf16((T, dm), dev); r1 = _f32((T,), dev)
            if dlt is None:
                L.rmsnorm_fwd(x, W[f"l{i}.san"], xn, None, r1)
            else:  # x = h_prev + mlp(h_prev), ref:307-308
                xs = _f32((T, dm), dev)
                L.rmsnorm_fwd(x, W[f"l{i}.san"], xn, None, r1, delta=dlt, h_out=xs)
                x = xs
            qkv = _bf16((T, QKV), dev)
            lora = None
            if not P.lora:
                tr = model.transformers
                L.gemm(xn, qkv_w, qkv, T, QKV, dm, epilogue=L.EPI_ROPE,
                       rope=dict(cos=tr.freqs_cos_t, sin=tr.freqs_sin_t, t_stride=int(tr.freqs_cos_t.shape[1]), cols=(H + G) * 64, row_len=2 * S))
            else:
                # LoRA on q and v (ref:263-271): q += 2 * B_q (A_q drop(x)), v likewise; rank padded 8 -> 64 with
                # zeros so the two skinny GEMMs use the regular tiles; the factor 2 = alpha/r is folded into B
                L.gemm(xn, qkv_w, qkv, T, QKV, dm)
                xq = xv = xn
                mq = mv = None
                if model.training:  # nn.Dropout(0.1) on the LoRA input, independent masks for q and v
                    xn32 = _f32((T, dm), dev)
                    L.rmsnorm_fwd(x, W[f"l{i}.san"], None, xn32, None)
                    mq, mv = model._lora_dropout(i, "q", (B, 2 * S, dm), dev), model._lora_dropout(i, "v", (B, 2 * S, dm), dev)
                    xq = _bf16((T, dm), dev); L.mul_to_bf16(xn32, mq, xq)
                    xv = _bf16((T, dm), dev); L.mul_to_bf16(xn32, mv, xv)
                    del xn32
                Aq, Bq2, Av, Bv2 = lw[i]
                aq = _bf16((T, 64), dev); L.gemm(xq, Aq, aq, T, 64, dm)
                L.gemm(aq, Bq2, qkv, T, H * 64, 64, ldc=QKV, epilogue=L.EPI_BF16_ACCUM)
                av = _bf16((T, 64), dev); L.gemm(xv, Av, av, T, 64, dm)
                L.gemm(av, Bv2, qkv[:, (H + G) * 64:], T, G * 64, 64, ldc=QKV, epilogue=L.EPI_BF16_ACCUM)
                L.rope(qkv, H + G, 2 * S, cos, sin)
                lora = (aq, av, xq, xv, mq, mv)
            ao = _bf16((T, H * 64), dev); lse = _f32((H, T), dev)
            at = L.attn_args(qkv, H, G, B, 2 * S, tm_tok, seg_s, seg_e, mode=_attn_mode(cfg))
            L.attn_fwd(at, ao, lse)
            dl1 = _bf16((T, dm), dev)
            L.gemm(ao, o_w, dl1, T, dm, H * 64)
            h = _f32((T, dm), dev); hn = _bf16((T, dm), dev); r2 = _f32((T,), dev)
            L.rmsnorm_fwd(x, W[f"l{i}.mln"], hn, None, r2, delta=dl1, h_out=h)  # h = x + attn(...), ref:304-306
            gu = _bf16((T, 2 * I), dev); act = _bf16((T, I), dev)
            L.gemm(hn, w13, gu, T, 2 * I, dm, epilogue=L.EPI_SWIGLU, aux=act, ldaux=I)
            dlt = dl1  # reuse the buffer for the mlp output
            L.gemm(act, w2, dlt, T, dm, I)
            layers.append((x, xn, r1, qkv, ao, lse, h, hn, r2, gu, act, lora))
            x = h
        hidden = _f32((T, dm), dev); rF = _f32((T,), dev)
        if dlt is None:
            L.rmsnorm_fwd(x, W["norm"], None, hidden, rF)
        else:
            xs = _f32((T, dm), dev)
            L.rmsnorm_fwd(x, W["norm"], None, hidden, rF, delta=dlt, h_out=xs)
            x = xs
        sv["layers"], sv["xL"], sv["rF"] = layers, x, rF
        if model._taps is not None:
            model._taps.update(x0=layers[0][0] if nl else x, hidden=hidden,
                               **{f"layer{i}": (layers[i + 1][0] if i + 1 < nl else x) for i in range(nl)})

        # ---- heads (ref:499-529) ----
        cap = cfg["mask_topk"] * B
        VP = comm.world() if cfg.get("vocab_parallel") and comm.world() > 1 else 0
        vrank = comm.rank()
        out = torch.zeros(16, dtype=torch.float32, device=dev)  # [4 losses | 2x4 rating moments(+den)]
        w0_16 = _bf16((dm, dm), dev); L.cast_bf16(W["rh.w0"], w0_16)
        sv["w0_16"] = w0_16
        heads = []
        for m in ALL_MEDIUMS:
            lo, hi = (0, V0) if m == 0 else (V0, Vt)
            K_m = hi - lo
            # watch: fused logits + cross-entropy; [rows x catalog] never materialised
            sel, ws, lab, pos, cnt = L.select_rows(d[f"{m}.watch.weight"].reshape(-1), d[f"{m}.watch.label"].reshape(-1),
                                                   d[f"{m}.watch.position"].reshape(-1).to(torch.int32), cap)
            hs = _bf16((cap, dm), dev); L.gather_rows_bf16(hidden, sel, 0, hs)
            lse_r = _f32((cap,), dev); coef = _f32((cap,), dev); term = _f32((cap,), dev)
            vp = None
            if not VP:
                nb = L.ce_nblocks(K_m)
                pmax = _f32((nb, cap), dev); psum = _f32((nb, cap), dev); tgt = torch.zeros(cap, dtype=torch.float32, device=dev)
                # only the row blocks holding masked rows are computed: the count stays on the device
                L.gemm(hs, table16[lo:hi], None, cap, K_m, dm, epilogue=L.EPI_CE_PARTIAL,
                       ce=dict(pmax=pmax, psum=psum, tgt=tgt, target=pos), m_count=cnt)
                L.ce_finish(pmax, psum, nb, tgt, lab, ws, lse_r, coef, term, out[2 * m:2 * m + 1])
            else:
                # vocab-parallel catalog (SURVEY.md 8e / BASELINE config 4): this rank scores EVERY rank's
                # masked rows against its own slice of the medium's catalog; ranks then exchange
                # (max, sum exp, target logit) per row - 12 bytes instead of K_m logits
                R = VP * cap
                s0, s1 = lo + (K_m * vrank) // VP, lo + (K_m * (vrank + 1)) // VP
                hs_all = comm.all_gather_rows(hs)
                pos_loc = comm.all_gather_rows(pos) - (s0 - lo)  # targets outside [0, s1-s0) never match a column
                nb = L.ce_nblocks(s1 - s0)
                pmax = _f32((nb, R), dev); psum = _f32((nb, R), dev)
                st = torch.zeros((3, R), dtype=torch.float32, device=dev)
                L.gemm(hs_all, table16[s0:s1], None, R, s1 - s0, dm, epilogue=L.EPI_CE_PARTIAL,
                       ce=dict(pmax=pmax, psum=psum, tgt=st[2], target=pos_loc))
                L.ce_merge_partials(pmax, psum, nb, st[0], st[1])
                got = comm.all_to_all_rows(st.view(3, VP, cap).permute(1, 0, 2).contiguous())  # [shard, 3, my rows]
                L.ce_finish(got[:, 0].contiguous(), got[:, 1].contiguous(), VP, got[:, 2].sum(0), lab, ws, lse_r, coef, term,
                            out[2 * m:2 * m + 1])
                vp = dict(hs_all=hs_all, pos_loc=pos_loc, s0=s0, s1=s1)
            # rating: Linear -> GELU -> Linear(d,1), weighted MSE / moments
            rsel, rws, rlab, _rpos, rcnt = L.select_rows(d[f"{m}.rating.weight"].reshape(-1), d[f"{m}.rating.label"].reshape(-1),
                                                         d[f"{m}.rating.position"].reshape(-1).to(torch.int32), cap)
            rhs = _bf16((cap, dm), dev); L.gather_rows_bf16(hidden, rsel, 1, rhs)
            # only the row blocks that hold rating targets are computed (device-side limit); the rest stays zero
            h1 = torch.zeros((cap, dm), dtype=torch.bfloat16, device=dev); pre = torch.zeros((cap, dm), dtype=torch.bfloat16, device=dev)
            L.gemm(rhs, w0_16, h1, cap, dm, dm, epilogue=L.EPI_BIAS_GELU, bias=W["rh.b0"], aux=pre, m_count=rcnt)
            pred = _f32((cap,), dev)
            loss4 = out[4 + 4 * m: 8 + 4 * m]
            L.rating_tail(h1, W["rh.w2"], W["rh.b2"], rlab, rws, cfg["rating_mean"], pred, loss4)
            heads.append(dict(lo=lo, hi=hi, sel=sel, hs=hs, pos=pos, lse=lse_r, coef=coef, cnt=cnt, vp=vp, rsel=rsel, rhs=rhs, h1=h1, pre=pre,
                              pred=pred, rlab=rlab, rws=rws, rcnt=rcnt))
        out[1] = out[4]  # train MSE = first moment
        out[3] = out[8]
        sv["heads"], sv["out"] = heads, out
        sv["dims"] = (B, S, N, T, dm, H, G, I, nl, QKV, V0, Vt, Mp, cap)
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
        cfg = model.config
        P = model._params
        params = ctx.saved_tensors
        W = dict(zip(P.names, params))
        B, S, N, T, dm, H, G, I, nl, QKV, V0, Vt, Mp, cap = sv["dims"]
        dev = gout.device
        gout = gout.contiguous().float()
        G_ = {}  # name -> grad tensor
        need = dict(zip(P.names, ctx.needs_input_grad[3:]))  # finetune: only the LoRA matrices (ref:361-371)
        base = not P.lora          # the block / head / embedding weights train
        tab = not P.fused          # the item table trains (pre-train mode)

        dtable = None
        if tab:
            dtable = _f32((Vt + 1, dm), dev)
            dtable[Vt].zero_()  # rows [0,Vt) are fully written by the head GEMMs below
        dhid = torch.zeros((T, dm), dtype=torch.float32, device=dev)
        table16 = sv["table16"]
        # catalog columns per backward chunk of the cloze head: 16384 columns x the ~19 live row blocks = ~1200 tiles for the
        # logit-gradient GEMM (8 waves) and 384 tiles for d table; round 1's 4096-column chunks left every launch
        # wave-quantised (57 tiles on 148 SMs for d hidden) and cost 69 launches per step
        CH = 16384
        dw0 = _f32((dm, dm), dev); db0 = torch.zeros(dm, dtype=torch.float32, device=dev)
        dw2 = torch.zeros(dm, dtype=torch.float32, device=dev); db2 = torch.zeros(1, dtype=torch.float32, device=dev)
        tmp_dw2 = _f32((dm,), dev); tmp_db2 = _f32((1,), dev); tmp_db0 = _f32((dm,), dev)
        for m, hd in enumerate(sv["heads"]):
            lo, hi = hd["lo"], hd["hi"]
            K_m = hi - lo
            # ---- watch head backward: recompute logits chunk-wise, never store [rows x catalog] ----
            coef = hd["coef"].clone(); L.scale_by_scalar(coef, gout[2 * m])
            if hd["vp"] is None:
                rows_n, hs_rows, tgt_rows, lse_rows, coef_rows = cap, hd["hs"], hd["pos"], hd["lse"], coef
                c_lo, c_hi, lim = lo, hi, hd["cnt"]
            else:  # vocab-parallel: all ranks' rows against this rank's catalog slice
                vp = hd["vp"]
                rows_n, hs_rows, tgt_rows = vp["hs_all"].shape[0], vp["hs_all"], vp["pos_loc"]
                lse_rows, coef_rows = comm.all_gather_rows(hd["lse"]), comm.all_gather_rows(coef)
                c_lo, c_hi, lim = vp["s0"], vp["s1"], None
                if tab:
                    dtable[lo:c_lo].zero_(); dtable[c_hi:hi].zero_()  # other slices: their owners write them
            dhs = _f32((rows_n, dm), dev)
            dlog = _bf16((rows_n, CH), dev)
            for c0 in range(0, c_hi - c_lo, CH):
                cn = min(CH, c_hi - c_lo - c0)
                items = table16[c_lo + c0: c_lo + c0 + cn]
                L.gemm(hs_rows, items, dlog, rows_n, cn, dm, ldc=CH, epilogue=L.EPI_CE_GRAD,
                       ce=dict(lse=lse_rows, coef=coef_rows, target=tgt_rows), n_offset=c0, m_count=lim)
                # d hidden rows += dlogits . items
                # (only ~2350 / ~1550 of the 8192 rows are live: 10 / 7 row blocks of 256 x 3 column blocks; five K splits.  Seven
                # (210 / 147 work units = 2.84 / 1.99 waves on 74 SM pairs instead of 2.03 / 1.42) measured no better: 572 vs 603 TFLOP/s)
                L.gemm(dlog, items, dhs, rows_n, dm, cn, lda=CH, b_kmajor=False, ldb=dm,
                       epilogue=L.EPI_F32 if c0 == 0 else L.EPI_F32_ACCUM, m_count=lim, split_k=5 if cn >= 4096 else 0)
                # d table rows = dlogits^T . hidden rows
                if tab:
                    L.gemm(dlog, hs_rows, dtable[c_lo + c0: c_lo + c0 + cn], cn, dm, rows_n, a_kmajor=False, lda=CH,
                           b_kmajor=False, ldb=dm, epilogue=L.EPI_F32, split_k=-1, k_count=lim)
            if hd["vp"] is not None:
                dhs = comm.reduce_scatter_rows(dhs)  # sum of the shards' partial d(hidden rows), my rows only
            L.scatter_rows_f32(dhs, hd["sel"], 0, dhid)
            # ---- rating head backward ----
            dpre = _bf16((cap, dm), dev); dp = _f32((cap + 32 * (dm + 1),), dev)
            loss4 = sv["out"][4 + 4 * m: 8 + 4 * m]
            L.rating_tail_bwd(hd["pre"], hd["h1"], W["rh.w2"], hd["rlab"], hd["rws"], cfg["rating_mean"], hd["pred"], loss4,
                              gout[2 * m + 1], dpre, dp, tmp_dw2, tmp_db2)
            if base:
                dw2 += tmp_dw2; db2 += tmp_db2
                L.gemm(dpre, hd["rhs"], dw0, dm, dm, cap, a_kmajor=False, lda=dm, b_kmajor=False, ldb=dm,
                       epilogue=L.EPI_F32 if m == 0 else L.EPI_F32_ACCUM, split_k=-1, k_count=hd["rcnt"])
                L.colsum(dpre, tmp_db0); db0 += tmp_db0
            drhs = _f32((cap, dm), dev)
            L.gemm(dpre, sv["w0_16"], drhs, cap, dm, dm, b_kmajor=False, ldb=dm, epilogue=L.EPI_F32, m_count=hd["rcnt"])
            L.scatter_rows_f32(drhs, hd["rsel"], 1, dhid)
        G_["rh.w0"], G_["rh.b0"], G_["rh.w2"], G_["rh.b2"] = dw0, db0, dw2.reshape(1, dm), db2

        # ---- final norm ----
        dx = _f32((T, dm), dev); dx16 = _bf16((T, dm), dev)
        dn = _f32((dm,), dev)
        L.rmsnorm_bwd(sv["xL"], W["norm"], sv["rF"], None, dhid, None, dx, dx16, dn)
        G_["norm"] = dn
        del dhid
        tm_tok, seg_s, seg_e = sv["seg"]
        cos, sin = model.transformers.freqs_cos, model.transformers.freqs_sin

        for i in reversed(range(nl)):
            x, xn, r1, qkv, ao, lse, h, hn, r2, gu, act, lora = sv["layers"][i]
            qkv_w, o_w, w13, w2, o_t = sv["wq"][i]
            # x_out = h + act . W2^T.  d(act) = dx . W2 and the SwiGLU backward in one kernel: the GEMM's epilogue reads the
            # gate / up tiles of `gu` and writes d(gate | up); d(act) itself never reaches HBM
            dgu = _bf16((T, 2 * I), dev)
            L.gemm(dx16, w2, dgu, T, I, dm, b_kmajor=False, ldb=I, ldc=2 * I, epilogue=L.EPI_SWIGLU_BWD, aux=gu, ldaux=2 * I)
            if base:
                g_w2 = _f32((dm, I), dev)
                L.gemm(dx16, act, g_w2, dm, I, T, a_kmajor=False, lda=dm, b_kmajor=False, ldb=I, epilogue=L.EPI_F32, split_k=-1)
            dhn = _bf16((T, dm), dev)
            L.gemm(dgu, w13, dhn, T, dm, 2 * I, b_kmajor=False, ldb=dm)
            if base:
                g_w13 = _f32((2 * I, dm), dev)
                L.gemm(dgu, hn, g_w13, 2 * I, dm, T, a_kmajor=False, lda=2 * I, b_kmajor=False, ldb=dm, epilogue=L.EPI_F32,
                       split_k=-1)
            del dgu
            dh = _f32((T, dm), dev); dh16 = _bf16((T, dm), dev); g_mln = _f32((dm,), dev)
            L.rmsnorm_bwd(h, W[f"l{i}.mln"], r2, dhn, None, dx, dh, dh16, g_mln)
            # h = x + ao . Wo^T
            dao = _bf16((T, H * 64), dev)
            # against a transposed copy of Wo (K-major B): the MN-major form of this 768-deep shape ran at 39 % of peak
            L.gemm(dh16, o_t, dao, T, H * 64, dm)
            if base:
                g_o = _f32((dm, H * 64), dev)
                L.gemm(dh16, ao, g_o, dm, H * 64, T, a_kmajor=False, lda=dm, b_kmajor=False, ldb=H * 64, epilogue=L.EPI_F32,
                       split_k=-1)
            dqkv = _bf16((T, QKV), dev); delta = _f32((H, T), dev)
            fused_unrope = (H // G) % 2 == 0  # the pair kernels un-rotate dq / dk in their final store
            at = L.attn_args(qkv, H, G, B, 2 * S, tm_tok, seg_s, seg_e, mode=_attn_mode(cfg), rope=(cos, sin) if fused_unrope else None)
            L.attn_bwd(at, ao, dao, lse, dqkv, delta)
            if not fused_unrope:
                L.rope(dqkv, H + G, 2 * S, cos, sin, inverse=True)
            dxn = dhn  # reuse
            if lora is not None:
                # q = q0 + (A_q xq) B2_q^T with B2 = 2B:  d(A x) = dq . B2,  dB = 2 dq^T (A x),  dA = d(Ax)^T xq
                aq, av, xq, xv, mq, mv = lora
                Aq, Bq2, Av, Bv2 = sv["lw"][i]
                d_aq = _bf16((T, 64), dev); d_av = _bf16((T, 64), dev)
                dv_view = dqkv[:, (H + G) * 64:]
                L.gemm(dqkv, Bq2, d_aq, T, 64, H * 64, lda=QKV, b_kmajor=False, ldb=64)
                L.gemm(dv_view, Bv2, d_av, T, 64, G * 64, lda=QKV, b_kmajor=False, ldb=64)
                gBq = _f32((H * 64, 64), dev); gBv = _f32((G * 64, 64), dev); gAq = _f32((64, dm), dev); gAv = _f32((64, dm), dev)
                L.gemm(dqkv, aq, gBq, H * 64, 64, T, a_kmajor=False, lda=QKV, b_kmajor=False, ldb=64, epilogue=L.EPI_F32, split_k=-1)
                L.gemm(dv_view, av, gBv, G * 64, 64, T, a_kmajor=False, lda=QKV, b_kmajor=False, ldb=64, epilogue=L.EPI_F32, split_k=-1)
                L.gemm(d_aq, xq, gAq, 64, dm, T, a_kmajor=False, lda=64, b_kmajor=False, ldb=dm, epilogue=L.EPI_F32, split_k=-1)
                L.gemm(d_av, xv, gAv, 64, dm, T, a_kmajor=False, lda=64, b_kmajor=False, ldb=dm, epilogue=L.EPI_F32, split_k=-1)
                G_[f"l{i}.qB"], G_[f"l{i}.vB"] = 2.0 * gBq[:, :8], 2.0 * gBv[:, :8]
                G_[f"l{i}.qA"], G_[f"l{i}.vA"] = gAq[:8], gAv[:8]
                if i == 0 and not tab:
                    sv["layers"][i] = None
                    break  # nothing below the first block trains (ref:361-371)
            L.gemm(dqkv, qkv_w, dxn, T, dm, QKV, b_kmajor=False, ldb=dm)
            if lora is not None:  # d(norm output) also flows through the LoRA branches (and their dropout masks)
                if mq is None:
                    L.gemm(d_aq, Aq, dxn, T, dm, 64, b_kmajor=False, ldb=dm, epilogue=L.EPI_BF16_ACCUM)
                    L.gemm(d_av, Av, dxn, T, dm, 64, b_kmajor=False, ldb=dm, epilogue=L.EPI_BF16_ACCUM)
                else:
                    tmp = _bf16((T, dm), dev)
                    L.gemm(d_aq, Aq, tmp, T, dm, 64, b_kmajor=False, ldb=dm); L.fma_mask_bf16(dxn, tmp, mq)
                    L.gemm(d_av, Av, tmp, T, dm, 64, b_kmajor=False, ldb=dm); L.fma_mask_bf16(dxn, tmp, mv)
            if base:
                g_qkv = _f32((QKV, dm), dev)
                L.gemm(dqkv, xn, g_qkv, QKV, dm, T, a_kmajor=False, lda=QKV, b_kmajor=False, ldb=dm, epilogue=L.EPI_F32, split_k=-1)
            g_san = _f32((dm,), dev)
            L.rmsnorm_bwd(x, W[f"l{i}.san"], r1, dxn, None, dh, dx, dx16, g_san)
            if base:
                G_[f"l{i}.q"], G_[f"l{i}.k"], G_[f"l{i}.v"] = g_qkv[: H * 64], g_qkv[H * 64:(H + G) * 64], g_qkv[(H + G) * 64:]
                g13 = g_w13.view(I // 64, 2, 64, dm)  # rows are in the interleaved gate/up block order
                G_[f"l{i}.o"], G_[f"l{i}.w1"], G_[f"l{i}.w3"], G_[f"l{i}.w2"] = g_o, g13[:, 0].reshape(I, dm), g13[:, 1].reshape(I, dm), g_w2
                G_[f"l{i}.san"], G_[f"l{i}.mln"] = g_san, g_mln
            sv["layers"][i] = None  # free this layer's activations

        if not base:  # finetune: embeddings, heads and block weights are frozen
            ctx.sv = None
            return (None, None, None, *[(G_[n].contiguous() if nd else None) for n, nd in need.items()])

        # ---- token embeddings ----
        L.scatter_add_sorted(dx, 2 * dm, sv["mid"], dm, dtable)  # item rows of the interleaved stream
        ae = {k: torch.zeros_like(W[k]) for k in ("ae.w", "ae.b", "ae.gender", "ae.source", "ae.status", "ae.cos", "ae.sin")}
        L.action_embed_bwd(sv["aa"], sv["feats"], dx, ae["ae.w"], ae["ae.b"], ae["ae.gender"], ae["ae.source"], ae["ae.status"],
                           ae["ae.cos"], ae["ae.sin"])
        G_.update(ae)
        # fused table -> id rows, projection weight and bias (ref:139-145, 165-169)
        G_["ie.id"] = dtable
        dpb = _f32((dm,), dev); L.colsum(dtable, dpb)
        G_["ie.pb"] = dpb
        dt16 = _bf16((Vt + 1, dm), dev); L.cast_bf16(dtable, dt16)
        meta16, _ = model.item_embedding.meta_bf16()
        Mraw = cfg["metadata_emb_size"]
        dpw = _f32((dm, Mraw), dev)
        L.gemm(dt16, meta16, dpw, dm, Mraw, Vt + 1, a_kmajor=False, lda=dm, b_kmajor=False, ldb=Mp, epilogue=L.EPI_F32, split_k=-1)
        G_["ie.pw"] = dpw
        ctx.sv = None
        grads = []
        for n, needs in zip(P.names, ctx.needs_input_grad[3:]):
            grads.append(G_[n].contiguous() if needs else None)
        return (None, None, None, *grads)


# ------------------------------------------------------------------------------
@torch.no_grad()
def _inference_hidden(model, d):
    """ref:464-491 + 335-343 without autograd state: `d` holds [U, S'] arrays incl. optional
    `rope_input_pos` (ref:470-476) and arbitrary `token_mask_ids` (embed.py:122-125).  Returns the
    final-norm output [U, 2S', d] fp32.  Same kernels as the train forward; nothing is kept."""
    cfg, P = model.config, model._params
    W = dict(zip(P.names, P.tensors))
    dev = W["ae.w"].device
    if dev.type != "cuda":
        raise L.GraftError("the inference forward has no CPU path: move the model and batch to a CUDA device")
    B, S = d["userid"].shape
    N, T = B * S, 2 * B * S
    dm, H, G, I = cfg["embed_dim"], cfg["num_heads"], cfg["num_kv_heads"], cfg["intermediate_dim"]
    assert dm // H == 64 and I % 64 == 0 and (2 * S) % 128 == 0, "head_dim 64, FFN % 64, row length % 64 positions"
    QKV = (H + 2 * G) * 64
    d = _normalise_batch({k: (v if v.is_contiguous() else v.contiguous()) for k, v in d.items()})
    wts = model._inference_weights()
    table = wts["table"]
    x = _f32((T, dm), dev)
    L.item_gather(table, d["matchedid"].reshape(-1), x)
    L.action_embed(model.action_embedding.kernel_args(d, W, N), x, _f32((N, 32), dev))
    tm_tok = torch.empty(T, dtype=torch.int32, device=dev); L.interleave_i32(d["token_mask_ids"].reshape(-1), tm_tok)
    seg_s, seg_e = L.token_segments(d["userid"])
    cos, sin = model.transformers.freqs_cos, model.transformers.freqs_sin
    rope = dict(cos=cos, sin=sin, cols=(H + G) * 64, row_len=2 * S)
    if "rope_input_pos" in d:  # ref:470-474: token positions 2p (item) and 2p+1 (action)
        rp = d["rope_input_pos"].reshape(-1).to(torch.int32)
        rope["pos"] = torch.stack((2 * rp, 2 * rp + 1), dim=1).reshape(-1).contiguous()
        assert int(cos.shape[0]) >= 2, "rope table missing"
    else:
        assert 2 * S <= cos.shape[0], "row longer than the RoPE table (2 * max_sequence_length)"
    xn = _bf16((T, dm), dev); qkv = _bf16((T, QKV), dev); ao = _bf16((T, H * 64), dev); lse = _f32((H, T), dev)
    gu = _bf16((T, 2 * I), dev); act = _bf16((T, I), dev)
    dl = _bf16((T, dm), dev)
    h = _f32((T, dm), dev)
    first = True
    for i in range(cfg["num_layers"]):
        qkv_w, o_w, w13, w2 = wts["layers"][i]
        if first:
            L.rmsnorm_fwd(x, W[f"l{i}.san"], xn, None, None)
        else:
            L.rmsnorm_fwd(h, W[f"l{i}.san"], xn, None, None, delta=dl, h_out=x)  # x = h + mlp(h)
        first = False
        L.gemm(xn, qkv_w, qkv, T, QKV, dm, epilogue=L.EPI_ROPE, rope=rope)
        L.attn_fwd(L.attn_args(qkv, H, G, B, 2 * S, tm_tok, seg_s, seg_e, mode=_attn_mode(cfg)), ao, lse)
        L.gemm(ao, o_w, dl, T, dm, H * 64)
        L.rmsnorm_fwd(x, W[f"l{i}.mln"], xn, None, None, delta=dl, h_out=h)      # h = x + attn(x)
        L.gemm(xn, w13, gu, T, 2 * I, dm, epilogue=L.EPI_SWIGLU, aux=act, ldaux=I)
        L.gemm(act, w2, dl, T, dm, I)
    hidden = _f32((T, dm), dev)
    if first:
        L.rmsnorm_fwd(x, W["norm"], None, hidden, None)
    else:
        L.rmsnorm_fwd(h, W["norm"], None, hidden, None, delta=dl, h_out=x)
    return hidden.view(B, 2 * S, dm)


class RecommenderModel(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.action_embedding = ActionEmbedding(config)
        self.item_embedding = ItemEmbedding(config)
        self.transformers = Transformer(config)
        self.watch_head = DualItemEmbedding(self.item_embedding, config["vocab_sizes"]["0_matchedid"])
        self.rating_head = nn.Sequential(
            nn.Linear(config["embed_dim"], config["embed_dim"]), nn.GELU(), nn.Linear(config["embed_dim"], 1))
        self.apply(init_weights)
        if config["finetune"]:  # ref:361-371: everything frozen except the LoRA matrices
            for layer in (self.action_embedding, self.item_embedding, self.watch_head, self.rating_head):
                for p in layer.parameters():
                    p.requires_grad = False
            for name, p in self.transformers.named_parameters():
                p.requires_grad = "lora_" in name
        self._params_cache = None
        self._fused_cache = None
        self.lora_mask = None  # tests may inject f(layer, which, shape, device) -> dropout mask (0 or 1/0.9)
        self._taps = None
        self.mask_rand = None  # tests may inject f(shape, device) -> randval
        self._infer_cache = None
        if config["forward"] == "train":
            self.forward = self.train_forward
        elif config["forward"] == "inference":
            self.forward = self.inference_forward
        else:
            raise AssertionError(config["forward"])

    @property
    def _params(self):
        """Flat view of the step's tensors; rebuilt when ItemEmbedding.fuse() changed the tree."""
        c = self._params_cache
        if c is None or c.fused != self.item_embedding.fused:
            c = self._params_cache = _Params(self)
        return c

    # ref:379-389, with the matrix passed in or read from <datadir>/media_embeddings.h5
    def load_pretrained_embeddings(self, datadir=None, metadata=None):
        if metadata is None:
            import h5py  # same dependency as the reference
            with h5py.File(f"{datadir}/media_embeddings.h5") as f:
                metadata = f["metadata"][:]
        w = self.item_embedding.metadata_embedding.embedding.weight
        mt = torch.as_tensor(metadata, dtype=w.dtype, device=w.device)
        assert mt.shape[0] + 1 == w.shape[0] and mt.shape[1] == w.shape[1]
        with torch.no_grad():
            w.zero_()
            w[: mt.shape[0]].copy_(mt)
        self.item_embedding._meta_cache = None
        self.item_embedding._table_cache = None

    def _fused_tables(self):
        """finetune: the fused item table (fp32 buffer, ref:120-133) and its bf16 copy for the tied head."""
        t = self.item_embedding.fused_embedding
        key = (t.data_ptr(), t._version, str(t.device))
        if self._fused_cache is None or self._fused_cache[0] != key:
            t16 = torch.empty(t.shape, dtype=torch.bfloat16, device=t.device)
            L.cast_bf16(t, t16)
            self._fused_cache = (key, t16)
        return t, self._fused_cache[1]

    def _lora_dropout(self, layer, which, shape, dev):
        """nn.Dropout(0.1) mask of the LoRA input (ref:247, 263-271), as a multiplier 0 or 1/0.9.  Drawn by
        torch on a ones tensor of the reference's shape and dtype, in the reference's order (q then v,
        block by block), so the generator is advanced exactly as the reference advances it."""
        if self.lora_mask is not None:
            return self.lora_mask(layer, which, shape, dev).reshape(-1).contiguous()
        return torch.nn.functional.dropout(torch.ones(shape, dtype=torch.float32, device=dev), 0.1, True).reshape(-1)

    def _inference_weights(self):
        """bf16 operand copies of the block weights and the fused item table (= ItemEmbedding.fuse(),
        ref:120-133), rebuilt only when a parameter changed (inference weights are static)."""
        P = self._params
        key = tuple((t.data_ptr(), t._version) for t in P.tensors) + (str(P.tensors[0].device), ItemEmbedding._weights_version)
        if self._infer_cache is not None and self._infer_cache[0] == key:
            return self._infer_cache[1]
        cfg = self.config
        W = dict(zip(P.names, P.tensors))
        dev = W["ae.w"].device
        dm, H, G, I = cfg["embed_dim"], cfg["num_heads"], cfg["num_kv_heads"], cfg["intermediate_dim"]
        QKV = (H + 2 * G) * 64
        layers = []
        for i in range(cfg["num_layers"]):
            wq_, wv_ = W[f"l{i}.q"].detach(), W[f"l{i}.v"].detach()
            if P.lora:
                # inference has no dropout, so the LoRA branch (ref:263-271) is a rank-8 update of the weight:
                # W + (alpha/r) B A, merged once in fp32 (the reference adds the two bf16 GEMM outputs instead;
                # the difference is bf16 rounding noise, inside the 1e-2 tolerance)
                wq_ = wq_ + 2.0 * (W[f"l{i}.qB"].detach() @ W[f"l{i}.qA"].detach())
                wv_ = wv_ + 2.0 * (W[f"l{i}.vB"].detach() @ W[f"l{i}.vA"].detach())
            qkv_w, o_w, w13, w2, _ = _block_weights_bf16(wq_, W[f"l{i}.k"], wv_, W[f"l{i}.o"], W[f"l{i}.w1"], W[f"l{i}.w3"],
                                                        W[f"l{i}.w2"], H, G, dm, I, dev)
            layers.append((qkv_w, o_w, w13, w2))
        table = self.item_embedding.table()
        w0_16 = _bf16((dm, dm), dev); L.cast_bf16(W["rh.w0"], w0_16)
        out = dict(layers=layers, table=table, w0_16=w0_16)
        self._infer_cache = (key, out)
        return out

    def fused_item_table(self):
        """ItemEmbedding.fuse() (ref:120-133): id rows + projected metadata rows, [V+1, d] fp32.
        What register.py:14-38 exports and what the tied head scores against."""
        return self._inference_weights()["table"]

    @torch.no_grad()
    def retrieval_scores(self, user_embedding, medium, idxs=None):
        """Finetune/embed.jl:86-90 compute_retrieval: p = softmax(W_medium . u)[idxs] for a batch of user embeddings
        ([U, d]: the final-norm state inference_forward(d, "retrieval") returns at the user's masked query token).  W is
        the medium's slice of the fused item table (what register.py exports as "<medium>.watch.weight").  The softmax
        statistics come from the logit GEMM's epilogue; returns [U, len(idxs) or K_medium] fp32."""
        tab = self.item_embedding.table()
        _need_cuda(tab, "retrieval_scores")
        V0 = self.config["vocab_sizes"]["0_matchedid"]
        lo, hi = (0, V0) if medium == 0 else (V0, self.item_embedding.vocab_size)
        K, dm = hi - lo, tab.shape[1]
        u = user_embedding.reshape(-1, dm).to(tab.device)
        U = u.shape[0]
        u16 = u.to(torch.bfloat16).contiguous()
        items16 = _bf16((K, dm), tab.device); L.cast_bf16(tab[lo:hi], items16)
        ldc = (K + 7) // 8 * 8
        logits = _bf16((U, ldc), tab.device)
        L.gemm(u16, items16, logits, U, K, dm, ldc=ldc)
        nb = L.ce_nblocks(K)
        pmax = _f32((nb, U), tab.device); psum = _f32((nb, U), tab.device)
        tgt = _f32((U,), tab.device); pos = torch.zeros(U, dtype=torch.int32, device=tab.device)
        L.gemm(u16, items16, None, U, K, dm, epilogue=L.EPI_CE_PARTIAL, ce=dict(pmax=pmax, psum=psum, tgt=tgt, target=pos))
        rmax = _f32((U,), tab.device); rsum = _f32((U,), tab.device)
        L.ce_merge_partials(pmax, psum, nb, rmax, rsum)
        if idxs is not None:
            idxs = torch.as_tensor(idxs, device=tab.device).to(torch.int32).contiguous()
        out = _f32((U, K if idxs is None else idxs.numel()), tab.device)
        L.softmax_gather(logits, rmax, rsum, idxs, out)
        return out

    @torch.compiler.disable
    def inference_forward(self, d, task):
        """ref:531-538: retrieval -> final-norm token states [U, 2S, d]; ranking -> rating head
        on every token [U, 2S, 1] (bf16, as the autocast reference returns it)."""
        hidden = _inference_hidden(self, d)
        if task == "retrieval":
            return hidden
        if task != "ranking":
            raise AssertionError(task)
        U, T2, dm = hidden.shape
        W = dict(zip(self._params.names, self._params.tensors))
        dev = hidden.device
        rows = U * T2
        x16 = _bf16((rows, dm), dev); L.cast_bf16(hidden.view(rows, dm), x16)
        h1 = _bf16((rows, dm), dev); pre = _bf16((rows, dm), dev)
        L.gemm(x16, self._inference_weights()["w0_16"], h1, rows, dm, dm, epilogue=L.EPI_BIAS_GELU, bias=W["rh.b0"], aux=pre)
        pred = _f32((rows,), dev)
        L.rating_pred(h1, W["rh.w2"], W["rh.b2"], pred)
        return pred.view(U, T2, 1).to(torch.bfloat16)

    def mask_tokens(self, d):
        """ref:417-462: same draw (torch.rand, so a seeded run masks the same positions as the reference) and the same
        in-place edits of `d`, applied by one graft_mask_tokens launch instead of ~30 boolean-index assignments."""
        shape, dev = d["userid"].shape, d["userid"].device
        for k in d:  # the kernel edits the arrays in place: they must be the dict's own dense storage
            if not d[k].is_contiguous():
                d[k] = d[k].contiguous()
        if self.config["finetune"]:  # ref:418-435: the held-out targets of the finetuned metric are the masked positions
            pick = self.config["finetune_metric"]
            assert pick in ALL_METRICS
            L.mask_tokens(d, None, 0.0, finetune_metric=pick)
        else:
            randval = self.mask_rand(shape, dev) if self.mask_rand is not None else torch.rand(shape, device=dev)
            L.mask_tokens(d, randval.to(torch.float32).contiguous(), self.config["mask_rate"])
        return d

    @torch.compiler.disable
    def train_forward(self, d, evaluate):
        S = self.config["max_sequence_length"]
        if not self.config["finetune"]:
            for k in d:
                d[k] = d[k].reshape(-1, S)  # ref:495-496 (mutates the caller's dict, as the trainer expects)
        if not d["userid"].is_cuda or not self.rating_head[0].weight.is_cuda:
            raise L.GraftError("the cloze step has no CPU path: move the model and batch to a CUDA device")
        _normalise_batch(d)
        d = self.mask_tokens(d)
        for k in d:
            if not d[k].is_contiguous():
                d[k] = d[k].contiguous()
        if self.config.get("precision", "bf16") == "fp32":   # the step without autocast: fp32 kernels, parity 1e-4 (model_fp32.py)
            from .model_fp32 import ClozeStepF32
            res = ClozeStepF32.apply(self, d, bool(evaluate), *self._params.tensors)
        else:
            res = _ClozeStep.apply(self, d, bool(evaluate), *self._params.tensors)
        if evaluate:
            res, mom = res
            return [res[0], [mom[0, 0], mom[0, 1], mom[0, 2]], res[2], [mom[1, 0], mom[1, 1], mom[1, 2]]]
        return [res[0], res[1], res[2], res[3]]
