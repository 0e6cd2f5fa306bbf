"""Host mirror of the packing / batching either side of the step.

Mirrors, with the same names and argument meaning:
  * PretrainDataset.get_index_permutation / block_shuffle / window slicing
    (/root/reference/notebooks/Training/transformer.py:53-98)
  * the shard-by-rank rule (transformer.py:46-51)
  * concat (/root/reference/notebooks/Training/transformer.jl:144-163)
The index arithmetic runs in libgraft's plain-C host entry points
(graft_jagged_offsets_host, graft_block_permutation_host, graft_concat_pad_host:
the functions the Julia packer binds with ccall, integration/graft.jl), not in
Python loops.  No CUDA context is needed for anything in this file.
"""
from __future__ import annotations

import numpy as np

from . import lib as L


def jagged_offsets(userid):
    """Run boundaries of equal userid: offsets[0]=0 ... offsets[-1]=len (transformer.py:54)."""
    return L.jagged_offsets_host(userid)


def index_permutation(userid, permutation=np.random.permutation):
    """transformer.py:53-68: permutes whole user blocks; `permutation(n)` supplies the block order."""
    off = jagged_offsets(userid)
    order = np.asarray(permutation(len(off) - 1), dtype=np.int64)
    return L.block_permutation_host(off, order)


def block_shuffle(d, permutation=np.random.permutation):
    """transformer.py:70-73 (in place on the dict of flat arrays)."""
    p = index_permutation(d["userid"], permutation)
    for k in d:
        d[k] = d[k][p]
    return d


def windows(d, tokens_per_batch):
    """transformer.py:90-98: consecutive windows of tokens_per_batch positions."""
    n = len(d["userid"])
    assert n % tokens_per_batch == 0, "file length must be a multiple of the batch (transformer.py:91)"
    for i in range(0, n, tokens_per_batch):
        yield {k: v[i: i + tokens_per_batch] for k, v in d.items()}


def shards_for_rank(shards, local_rank, local_world_size):
    """transformer.py:46-51: shard directory i belongs to rank i % world."""
    shards = sorted(shards)
    assert len(shards) % local_world_size == 0
    return [s for i, s in enumerate(shards) if i % local_world_size == local_rank]


def concat(per_user, batch_size):
    """transformer.jl:144-163: every key back to back over users, zero padded to a multiple of batch_size."""
    n = sum(len(u["userid"]) for u in per_user)
    total = n + (-n) % batch_size
    return {k: L.concat_pad_host([u[k] for u in per_user], total) for k in per_user[0]}


# ------------------------------------------------------------------------------
# per-user featuriser: transformer.jl:79-142 `get_data` (+ history_tools.jl:37-75)
# ------------------------------------------------------------------------------
PLANNED_STATUS = 5                        # transformer.jl:18
PACK_METRICS = ("watch", "rating", "status")  # transformer.jl:17


def _columns(items):
    """List of msgpack item dicts -> struct of arrays (None -> NaN for the history_* fields)."""
    n = len(items)
    col = lambda k, dt: np.fromiter((x[k] for x in items), dtype=dt, count=n)  # noqa: E731
    opt = lambda k: np.fromiter((np.nan if x[k] is None else x[k] for x in items), dtype=np.float64, count=n)  # noqa: E731
    return dict(medium=col("medium", np.int32), matchedid=col("matchedid", np.int32), ts=col("history_max_ts", np.float64),
                status=col("status", np.int32), rating=col("rating", np.float32), progress=col("progress", np.float32),
                h_status=opt("history_status"), h_rating=opt("history_rating"))


def _tokenize_project(items):
    """`tokenize!` (collapse runs of one (medium, matchedid): first event's fields, last event's status / rating /
    progress) then `project!` (drop events that change neither status nor rating), history_tools.jl:37-75, on columns."""
    c = _columns(items)
    n0 = len(c["medium"])
    if n0:
        key = c["medium"].astype(np.int64) << 32 | c["matchedid"].astype(np.int64)
        start = np.flatnonzero(np.r_[True, key[1:] != key[:-1]])       # first event of each span
        last = np.r_[start[1:], n0] - 1                                 # last event of each span
    else:
        start = last = np.zeros(0, np.int64)
    out = dict(medium=c["medium"][start], matchedid=c["matchedid"][start], ts=c["ts"][start], h_status=c["h_status"][start],
               h_rating=c["h_rating"][start], status=c["status"][last], rating=c["rating"][last], progress=c["progress"][last])
    keep = ~((out["h_status"] == out["status"]) & (out["h_rating"] == out["rating"]))   # NaN (nothing) never equals: kept
    return {k: v[keep] for k, v in out.items()}


def _targets(c):
    """Watch / rating / status target hits of a token column set: transformer.jl:120-140 (= Finetune/transformer.jl:104-121)."""
    no_hist = np.isnan(c["h_status"])
    status, rating = c["status"], c["rating"]
    watch = ((status == 0) & no_hist) | ((status > PLANNED_STATUS) & (no_hist | ((c["h_status"] > 0) & (c["h_status"] <= PLANNED_STATUS))))
    rated = (rating > 0) & ~(rating == c["h_rating"])
    stat = (status > 0) & ~(status == c["h_status"])
    return {"watch": watch, "rating": rated, "status": stat}


def get_data_finetune(data, userid, num_items0, max_seq_len=1024, num_test_items=1):
    """Finetune packer row, /root/reference/notebooks/Finetune/transformer.jl:52-133: the user's tokenized + projected history
    (last max_seq_len - num_test_items tokens) followed by the held-out test items in one zero-padded row of max_seq_len
    positions; only the test items carry targets, and every (medium, metric) weight vector is normalised to sum 1.
    Returns the 27 arrays of length max_seq_len (one row of the [users, max_seq_len] matrices FinetuneDataset reads)."""
    assert len(data["test_items"]) <= num_test_items
    h = _tokenize_project(data["items"])
    keep = max_seq_len - num_test_items
    if len(h["medium"]) > keep:
        h = {k: v[len(v) - keep:] for k, v in h.items()}
    t = _columns(data["test_items"])
    nh, nt = len(h["medium"]), len(t["medium"])
    n = nh + nt
    cat = {k: np.concatenate([h[k], t[k]]) for k in h}
    u = data["user"]
    N = max_seq_len

    def row(values, dtype):
        out = np.zeros(N, dtype)
        out[:n] = values
        return out

    d = {"userid": row(userid, np.int32), "token_mask_ids": np.zeros(N, np.int32), "time": row(cat["ts"], np.float64),
         "gender": row(0 if u["gender"] is None else u["gender"] + 1, np.int32), "source": row(u["source"], np.int32),
         "matchedid": row(cat["matchedid"] + np.where(cat["medium"] == 1, num_items0, 0), np.int32),
         "status": row(cat["status"], np.int32), "rating": row(cat["rating"], np.float32), "progress": row(cat["progress"], np.float32)}
    hits = _targets(cat)
    is_test = np.r_[np.zeros(nh, bool), np.ones(nt, bool)]
    labels = {"watch": np.ones(n, np.float32), "rating": cat["rating"].astype(np.float32), "status": cat["status"].astype(np.float32)}
    d["token_mask_ids"] = row((hits["rating"] & is_test).astype(np.int32), np.int32)
    for m in (0, 1):
        for metric in PACK_METRICS:
            hit = hits[metric] & is_test & (cat["medium"] == m)
            w = hit.astype(np.float32)
            wsum = w.sum(dtype=np.float32)
            if wsum > 0:
                w = w / wsum
            d[f"{m}.{metric}.label"] = row(np.where(hit, labels[metric], 0), np.float32)
            d[f"{m}.{metric}.weight"] = row(w, np.float32)
            d[f"{m}.{metric}.position"] = row(np.where(hit, cat["matchedid"], 0), np.int32)
    return d


def get_data(data, userid, num_items0):
    """Columnar restatement of the Julia packer's per-user step: `tokenize!` (collapse runs of
    one (medium, matchedid): first event's fields, last event's status/rating/progress),
    `project!` (drop events that change neither status nor rating), then the 9 feature and 18
    target vectors with the reference's dtypes.  No per-item Python after the column build."""
    c = _columns(data["items"])
    n0 = len(c["medium"])
    if n0:
        key = c["medium"].astype(np.int64) << 32 | c["matchedid"].astype(np.int64)
        start = np.flatnonzero(np.r_[True, key[1:] != key[:-1]])       # first event of each span
        last = np.r_[start[1:], n0] - 1                                 # last event of each span
    else:
        start = last = np.zeros(0, np.int64)
    medium, mid, ts = c["medium"][start], c["matchedid"][start], c["ts"][start]
    h_status, h_rating = c["h_status"][start], c["h_rating"][start]
    status, rating, progress = c["status"][last], c["rating"][last], c["progress"][last]
    keep = ~((h_status == status) & (h_rating == rating))               # NaN (nothing) never equals: kept
    medium, mid, ts, h_status, h_rating = medium[keep], mid[keep], ts[keep], h_status[keep], h_rating[keep]
    status, rating, progress = status[keep], rating[keep], progress[keep]
    n = len(mid)
    u = data["user"]
    d = {
        "userid": np.full(n, userid, np.int32), "token_mask_ids": np.zeros(n, np.int32), "time": ts.astype(np.float64),
        "gender": np.full(n, 0 if u["gender"] is None else u["gender"] + 1, np.int32),
        "source": np.full(n, u["source"], np.int32),
        "matchedid": (mid + np.where(medium == 1, num_items0, 0)).astype(np.int32),
        "status": status.astype(np.int32), "rating": rating.astype(np.float32), "progress": progress.astype(np.float32),
    }
    no_hist = np.isnan(h_status)
    watch = ((status == 0) & no_hist) | ((status > PLANNED_STATUS) & (no_hist | ((h_status > 0) & (h_status <= PLANNED_STATUS))))
    rated = (rating > 0) & ~(rating == h_rating)
    stat = (status > 0) & ~(status == h_status)
    d["token_mask_ids"] = rated.astype(np.int32)
    labels = {"watch": np.ones(n, np.float32), "rating": rating.astype(np.float32), "status": status.astype(np.float32)}
    hits = {"watch": watch, "rating": rated, "status": stat}
    for m in (0, 1):
        for metric in PACK_METRICS:
            h = hits[metric] & (medium == m)
            d[f"{m}.{metric}.label"] = np.where(h, labels[metric], 0).astype(np.float32)
            d[f"{m}.{metric}.weight"] = h.astype(np.float32)
            d[f"{m}.{metric}.position"] = np.where(h, mid, 0).astype(np.int32)
    return d
