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
