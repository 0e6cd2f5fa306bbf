"""Packer rows of SURVEY.md section 8 (a1-a3): the per-user featuriser, concat and block shuffle.

Chain of evidence: the reference's own tokenize/project (embed.py:39-71, executed from
/root/reference by tests/golden/make_packer_golden.py) -> golden JSON -> pins
oracle/packer_oracle.py (a literal restatement of the Julia) -> checks the columnar
product implementation recommendersystem_b200/packer.py bit for bit."""
import json
import os

import numpy as np

from oracle import packer_oracle as po
from recommendersystem_b200 import packer, synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _golden():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "packer_golden.json")))


def test_oracle_tokenize_project_match_reference_outputs():
    g = _golden()
    for user, tok, proj in zip(g["users"], g["tokenized"], g["projected"]):
        assert po.tokenize(user["items"]) == tok
        assert po.project(po.tokenize(user["items"])) == proj


def test_get_data_bit_exact_against_oracle():
    g = _golden()
    keys = set(synthetic.all_keys())
    for i, user in enumerate(g["users"]):
        ref = po.get_data(user, i + 1, 40)
        got = packer.get_data(user, i + 1, 40)
        assert set(got) == set(ref) == keys
        for k in ref:
            assert got[k].dtype == ref[k].dtype == synthetic.key_dtype(k), k
            assert np.array_equal(got[k], ref[k]), (i, k)
        # every token is a change of status or rating; rating targets set the mask id (transformer.jl:129-134)
        assert np.array_equal(got["token_mask_ids"] > 0, (got["0.rating.weight"] + got["1.rating.weight"]) > 0)


def test_concat_pads_to_batch_multiple():
    g = _golden()
    per_user = [packer.get_data(u, i + 1, 40) for i, u in enumerate(g["users"])]
    per_user = [d for d in per_user if len(d["userid"])]
    out = packer.concat(per_user, 128)
    n = sum(len(d["userid"]) for d in per_user)
    assert len(out["userid"]) % 128 == 0 and len(out["userid"]) - n < 128
    for k in out:
        assert out[k].dtype == per_user[0][k].dtype
        assert np.array_equal(out[k][:n], np.concatenate([d[k] for d in per_user]))
        assert not out[k][n:].any()                         # zero pad: userid 0, all weights 0 (transformer.jl:144-163)
    off = packer.jagged_offsets(out["userid"])
    assert len(off) - 1 == len(per_user) + (1 if len(out["userid"]) > n else 0)


def _finetune_users():
    """Users of the golden file with their last 1-2 events split off as held-out test items (what the finetune ETL does)."""
    g = _golden()
    out = []
    for i, user in enumerate(g["users"]):
        items = list(user["items"])
        if len(items) < 3:
            continue
        n_test = 1
        out.append({"user": user["user"], "items": items[:-n_test], "test_items": items[-n_test:]})
    return out


def test_finetune_get_data_bit_exact_against_oracle():
    """Finetune/transformer.jl:52-133 (literal restatement in oracle/packer_oracle.py) vs the columnar product packer, incl.
    truncation to the last max_seq_len - 1 history tokens and the weight normalisation."""
    users = _finetune_users()
    assert users
    for seq in (1024, 8):   # 8: forces the history truncation branch
        for i, user in enumerate(users):
            ref = po.get_data_finetune(user, i + 1, 40, max_seq_len=seq)
            got = packer.get_data_finetune(user, i + 1, 40, max_seq_len=seq)
            assert set(got) == set(ref) == set(synthetic.all_keys())
            for k in ref:
                assert got[k].shape == (seq,) and got[k].dtype == ref[k].dtype == synthetic.key_dtype(k), k
                assert np.array_equal(got[k], ref[k]), (seq, i, k)
            for m in (0, 1):
                for metric in ("watch", "rating", "status"):
                    w = got[f"{m}.{metric}.weight"].sum()
                    assert w == 0 or abs(w - 1) < 1e-6


def test_finetune_dataset_batches(tmp_path):
    """FinetuneDataset (transformer.py:101-160): shard-by-rank, only users with a target of the finetuned medium, batches of
    [batch_size, max_seq_len]; training mode takes one partition of 4, shuffles and pads by re-drawing."""
    from recommendersystem_b200.loader import FinetuneDataset
    users = _finetune_users()
    seq = 16
    rows = [packer.get_data_finetune(u, i + 1, 40, max_seq_len=seq) for i, u in enumerate(users)] * 3
    mat = {k: np.stack([r[k] for r in rows]) for k in rows[0]}
    for shard in (1, 2):
        os.makedirs(tmp_path / str(shard))
        np.savez(tmp_path / str(shard) / "1.npz", **mat)
    medium = 1
    has = [i for i in range(len(rows)) if mat[f"{medium}.watch.weight"][i].sum() > 0 or mat[f"{medium}.rating.weight"][i].sum() > 0]
    ds = FinetuneDataset(str(tmp_path), 0, 2, 4, False, medium)
    assert len(ds.fns) == 1
    got = list(ds)
    assert sum(b["userid"].shape[0] for b in got) == len(has)
    seen = np.concatenate([b["userid"][:, 0].numpy() for b in got])
    assert np.array_equal(seen, mat["userid"][has, 0])      # evaluation order = file order
    for b in got:
        assert b["userid"].shape[1] == seq and set(b) == set(synthetic.all_keys())
    np.random.seed(0)
    tr = FinetuneDataset(str(tmp_path), 1, 2, 4, True, medium)
    epoch = list(tr)
    assert all(b["userid"].shape == (4, seq) for b in epoch)
    assert tr.partition[0] == 1
    n_part = len([x for i, x in enumerate(has) if i % 4 == 0])
    assert sum(b["userid"].shape[0] for b in epoch) == -(-n_part // 4) * 4
