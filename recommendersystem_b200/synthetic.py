"""Synthetic user histories in the reference's packed on-disk layout.

Produces exactly what the Julia packer writes and the trainer's dataset yields
(reference: notebooks/Training/transformer.jl:79-163 `get_data`/`concat`,
notebooks/Training/transformer.py:53-98 window slicing): 27 flat arrays whose
length is a multiple of rows*seq, users packed back to back, zero padding with
userid 0.  Distributions follow SURVEY.md section 8(d).

numpy only: this module is input plumbing for tests and bench.py.
"""
from __future__ import annotations

import numpy as np

MEDIUMS = (0, 1)
PACK_METRICS = ("watch", "rating", "status")  # transformer.jl:17

FEATURE_DTYPES = {
    "userid": np.int32,
    "token_mask_ids": np.int32,
    "time": np.float64,
    "gender": np.int32,
    "source": np.int32,
    "matchedid": np.int32,
    "status": np.int32,
    "rating": np.float32,
    "progress": np.float32,
}


def all_keys():
    keys = list(FEATURE_DTYPES)
    for m in MEDIUMS:
        for metric in PACK_METRICS:
            keys += [f"{m}.{metric}.label", f"{m}.{metric}.weight", f"{m}.{metric}.position"]
    return keys


def key_dtype(k):
    if k in FEATURE_DTYPES:
        return FEATURE_DTYPES[k]
    return np.int32 if k.endswith(".position") else np.float32


# --------------------------------------------------------------------------- #
# named configurations
# --------------------------------------------------------------------------- #
MIN_TS = 946684800.0
MAX_TS = 1.76e9


def make_config(name="C1", **over):
    base = {
        "max_sequence_length": 256,
        "vocab_sizes": {"0_matchedid": 4000, "1_matchedid": 20000, "status": 9, "gender": 4, "source": 4},
        "num_layers": 2,
        "num_heads": 4,
        "num_kv_heads": 2,
        "embed_dim": 256,
        "intermediate_dim": 704,
        "metadata_emb_size": 64,
        "min_ts": MIN_TS,
        "max_ts": MAX_TS,
        "rating_mean": 7.6287384,
        "rating_std": 1.778219,
        "forward": "train",
        "finetune": False,
        "learning_rate": 1e-4,
        "mask_rate": 0.1,
        "mask_topk": 32,
    }
    rows = 8
    if name == "C1":
        pass
    elif name == "C2":
        base.update(
            max_sequence_length=1024, num_layers=8, num_heads=12, num_kv_heads=6, embed_dim=768,
            intermediate_dim=2112, metadata_emb_size=6148, mask_topk=128,
        )
        base["vocab_sizes"] = {"0_matchedid": 60000, "1_matchedid": 40000, "status": 9, "gender": 4, "source": 4}
        rows = 64
    elif name in ("C3", "C4", "C4L", "C5"):
        # BASELINE configs[2..4] = the C2 model under another objective / sharding / input distribution:
        #   C3  HSTU-style rating model: point-wise SiLU attention, rating-only task weights (bench / tests pick the weights)
        #   C4  vocab-parallel catalog logits + cross-entropy (needs world > 1 to differ from C2)
        #   C4L the same with the 1,000,000-item catalog stress of SURVEY 8d (metadata width cut to 512 so that the frozen
        #       metadata table, which is not what this config stresses, stays at 2 GB)
        #   C5  daily-finetune-like input: short recent histories, Zipf 1.2 items (sampler_kwargs) - scatter-add stress
        base, rows = make_config("C2")
        if name == "C3":
            base["attention"] = "silu"
        if name in ("C4", "C4L"):
            base["vocab_parallel"] = True
        if name == "C4L":
            base["vocab_sizes"] = {"0_matchedid": 600000, "1_matchedid": 400000, "status": 9, "gender": 4, "source": 4}
            base["metadata_emb_size"] = 512
    elif name == "P":  # reference production shape, transformer.py:535-560
        base.update(
            max_sequence_length=1024, num_layers=8, num_heads=32, num_kv_heads=16, embed_dim=2048,
            intermediate_dim=5632, metadata_emb_size=6148, mask_topk=128,
        )
        base["vocab_sizes"] = {"0_matchedid": 60000, "1_matchedid": 40000, "status": 9, "gender": 4, "source": 4}
        rows = 64
    elif name == "tiny":  # golden-fixture scale
        base.update(
            max_sequence_length=64, num_layers=2, num_heads=2, num_kv_heads=1, embed_dim=128,
            intermediate_dim=192, metadata_emb_size=40, mask_topk=16,
        )
        base["vocab_sizes"] = {"0_matchedid": 150, "1_matchedid": 250, "status": 9, "gender": 4, "source": 4}
        rows = 4
    else:
        raise KeyError(name)
    base.update(over)
    return base, rows


def sampler_kwargs(name):
    """HistorySampler settings that belong to a named workload."""
    if name == "C5":  # histories of 5..64 positions packed back to back (~1900 users per 65,536-position batch), heavy duplicates
        return dict(len_median=30.0, len_sigma=1.0, len_clip=(5, 64), zipf_s=1.2)
    return {}


# --------------------------------------------------------------------------- #
# generator
# --------------------------------------------------------------------------- #
class HistorySampler:
    """Draws users and packs them like `get_data` + `concat`."""

    def __init__(self, config, seed=0, len_median=150.0, len_sigma=1.0, len_clip=(5, 4000), zipf_s=1.0):
        self.config = config
        self.rng = np.random.default_rng(seed)
        self.V0 = int(config["vocab_sizes"]["0_matchedid"])
        self.V1 = int(config["vocab_sizes"]["1_matchedid"])
        self.V = self.V0 + self.V1
        self.len_median, self.len_sigma, self.len_clip = len_median, len_sigma, len_clip
        # popularity is independent of id: the reference shuffles matchedid
        # (notebooks/Training/import_data.jl:133-136)
        ranks = np.arange(1, self.V + 1, dtype=np.float64)
        cdf = np.cumsum(ranks ** (-float(zipf_s)))
        self._cdf = cdf / cdf[-1]
        self._perm = self.rng.permutation(self.V).astype(np.int32)

    def user_lengths(self, n_users):
        ln = self.rng.lognormal(np.log(self.len_median), self.len_sigma, size=n_users)
        return np.clip(np.rint(ln), self.len_clip[0], self.len_clip[1]).astype(np.int64)

    def pack(self, n_users, pad_multiple):
        """Returns (dict of 27 flat arrays, offsets[int64 n_users+1])."""
        rng = self.rng
        lens = self.user_lengths(n_users)
        offsets = np.zeros(n_users + 1, dtype=np.int64)
        np.cumsum(lens, out=offsets[1:])
        n = int(offsets[-1])
        total = n + (-n) % int(pad_multiple)
        d = {k: np.zeros(total, dtype=key_dtype(k)) for k in all_keys()}

        seg = np.repeat(np.arange(n_users, dtype=np.int64), lens)
        d["userid"][:n] = (seg + 1).astype(np.int32)  # transformer.jl:225, 1-based
        # per-user sorted timestamps: sort (user, t) lexicographically
        t = rng.uniform(self.config["min_ts"], self.config["max_ts"], size=n)
        order = np.lexsort((t, seg))
        d["time"][:n] = t[order]
        d["gender"][:n] = rng.integers(0, 4, size=n_users)[seg]
        d["source"][:n] = rng.integers(0, 4, size=n_users)[seg]
        item = self._perm[np.searchsorted(self._cdf, rng.random(n), side="right").clip(0, self.V - 1)]
        d["matchedid"][:n] = item
        medium = (item >= self.V0).astype(np.int64)
        local = (item - medium * self.V0).astype(np.int32)
        status = rng.integers(0, 9, size=n).astype(np.int32)
        d["status"][:n] = status
        rating = np.clip(np.rint(rng.normal(7.6, 1.8, size=n)), 1, 10)
        rating[rng.random(n) < 0.4] = 0
        d["rating"][:n] = rating.astype(np.float32)
        d["progress"][:n] = rng.random(n).astype(np.float32)

        watch = rng.random(n) < 0.6
        rate = (rating > 0) & (rng.random(n) < 0.5)
        stat = (status > 0) & (rng.random(n) < 0.5)
        d["token_mask_ids"][:n] = rate.astype(np.int32)  # transformer.jl:130
        idx = np.arange(n)
        for m in MEDIUMS:
            sel = medium == m
            for metric, hit, lab in (("watch", watch, None), ("rating", rate, rating), ("status", stat, status)):
                w = idx[sel & hit]
                d[f"{m}.{metric}.weight"][w] = 1.0
                d[f"{m}.{metric}.label"][w] = 1.0 if lab is None else lab[w].astype(np.float32)
                d[f"{m}.{metric}.position"][w] = local[w]
        return d, offsets


def make_batches(config, rows, n_batches, seed=0, **sampler_kw):
    """Consecutive windows of rows*seq positions (transformer.py:91-98)."""
    S = int(config["max_sequence_length"])
    per = rows * S
    sampler = HistorySampler(config, seed=seed, **sampler_kw)
    mean_len = float(np.mean(sampler.user_lengths(4096)))
    sampler = HistorySampler(config, seed=seed, **sampler_kw)
    n_users = max(1, int(np.ceil(per * n_batches / mean_len)))
    d, _ = sampler.pack(n_users, per)
    while len(d["userid"]) < per * n_batches:  # unlucky draw: add users
        n_users = int(n_users * 1.25) + 1
        sampler = HistorySampler(config, seed=seed, **sampler_kw)
        d, _ = sampler.pack(n_users, per)
    return [{k: v[i * per:(i + 1) * per].copy() for k, v in d.items()} for i in range(n_batches)]


def make_metadata(config, seed=0):
    """Frozen metadata matrix (V, M) as `media_embeddings.h5["metadata"]` holds it."""
    rng = np.random.default_rng(seed + 7919)
    V = config["vocab_sizes"]["0_matchedid"] + config["vocab_sizes"]["1_matchedid"]
    return (rng.standard_normal((V, config["metadata_emb_size"]), dtype=np.float32) * np.float32(0.02))
