"""CPU-side checks: libgraft.so loads, exports every symbol include/graft.h
declares, and the host packing entry points agree with the oracle.  No GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import cloze_oracle as oc
from recommendersystem_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def built():
    if not os.path.exists(L.LIB_PATH):
        import __graft_entry__ as g
        g.build()


def test_exports_match_header():
    hdr = open(os.path.join(ROOT, "include", "graft.h")).read()
    names = sorted(set(re.findall(r"^(?:int|int64_t|const char\*)\s+(graft_\w+)\s*\(", hdr, flags=re.M)))
    assert len(names) >= 35
    lib = ctypes.CDLL(L.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.graft_version() >= 100


def test_jagged_offsets_host_matches_oracle():
    rng = np.random.default_rng(0)
    for n_users in (0, 1, 5, 300):
        lens = rng.integers(1, 40, n_users)
        u = np.repeat(rng.permutation(n_users).astype(np.int32) + 1, lens)
        u = np.concatenate([u, np.zeros(rng.integers(0, 9), np.int32)])
        assert np.array_equal(L.jagged_offsets_host(u), oc.jagged_offsets(u))
    u = np.array([4, 4, 2, 4, 4, 4], np.int32)  # a userid may recur non-contiguously: runs, not groups
    assert L.jagged_offsets_host(u).tolist() == [0, 2, 3, 6]


def test_block_permutation_and_concat_host():
    rng = np.random.default_rng(1)
    u = np.repeat(np.arange(1, 50, dtype=np.int32), rng.integers(1, 30, 49))
    off = L.jagged_offsets_host(u)
    order = rng.permutation(len(off) - 1)
    assert np.array_equal(L.block_permutation_host(off, order), oc.block_permutation(off, order))
    with pytest.raises(L.GraftError):
        L.block_permutation_host(off, np.array([len(off) + 3] * (len(off) - 1)))
    pieces = [rng.standard_normal(int(k)).astype(np.float32) for k in rng.integers(0, 20, 7)]
    n = sum(len(p) for p in pieces)
    total = n + (-n) % 64
    got = L.concat_pad_host(pieces, total)
    ref = oc.concat_pad([{"userid": p} for p in pieces], 64)["userid"]
    assert np.array_equal(got, ref)


def test_synthetic_pack_is_reference_layout():
    from recommendersystem_b200 import synthetic
    cfg, rows = synthetic.make_config("C1")
    b = synthetic.make_batches(cfg, rows, 2, seed=3)
    assert len(b) == 2 and sorted(b[0]) == sorted(synthetic.all_keys()) and len(b[0]) == 27
    for k, v in b[0].items():
        assert v.dtype == synthetic.key_dtype(k) and v.shape == (rows * cfg["max_sequence_length"],)
    off = L.jagged_offsets_host(np.concatenate([b[0]["userid"], b[1]["userid"]]))
    assert len(off) > 10
    V0 = cfg["vocab_sizes"]["0_matchedid"]
    m1 = b[0]["1.watch.weight"] > 0
    assert np.array_equal(b[0]["1.watch.position"][m1], b[0]["matchedid"][m1] - V0)


def test_no_cpu_fallback():
    import torch
    from recommendersystem_b200.model import RecommenderModel
    from recommendersystem_b200 import synthetic
    cfg, rows = synthetic.make_config("tiny")
    model = RecommenderModel(cfg)
    d = {k: torch.as_tensor(v) for k, v in synthetic.make_batches(cfg, rows, 1)[0].items()}
    with pytest.raises(L.GraftError):
        model(d, False)


def test_state_dict_surface(golden):
    from recommendersystem_b200.testing import model_from_params
    model = model_from_params(golden["config"], golden["param"], "cpu")
    keys = set(model.state_dict())
    assert {k for k in keys if not k.startswith("watch_head.")} == set(golden["param"])
    assert {"watch_head." + k for k in golden["param"] if k.startswith("item_embedding.")} <= keys
    frozen = [n for n, p in model.named_parameters() if not p.requires_grad]
    assert frozen == ["item_embedding.metadata_embedding.embedding.weight"]


def test_gemm_split_plan_host():
    """The automatic split-K plan (host-side cost model, gemm_sm100.cu: pick_auto) for the backward's dW shapes at C2
    on a 148-SM part, read back through graft_gemm_scratch_bytes (= splits * M * N * 4); no kernel is launched."""
    lib = L.load()

    def splits(M, N, K, epilogue=L.EPI_F32, split_k=-1):
        g = L.GemmArgs()
        g.M, g.N, g.K, g.epilogue, g.split_k = M, N, K, epilogue, split_k
        need = lib.graft_gemm_scratch_bytes(ctypes.byref(g))
        assert need % (M * N * 4) == 0
        return need // (M * N * 4)

    # 256 x 256 tiles on the 74 SM pairs of a 148-SM part (one-CTA 128 x 256 tiles on 148 SMs give the same counts)
    assert splits(768, 2112, 131072) == 8      # 27 pair tiles -> 216 work units = 2.9 rounds of 74 pairs
    assert splits(4224, 768, 131072) == 7      # 51 pair tiles -> 357 units = 4.8 rounds
    assert splits(1536, 768, 131072) == 4      # 18 pair tiles -> 72 units: one round
    assert splits(768, 768, 131072) == 8       # 9 pair tiles -> 72 units
    assert splits(768, 6148, 100001) == 0      # 128-wide tiles: 294 tiles = 2 full rounds, no split needed
    assert splits(768, 2112, 131072, split_k=0) == 0 and splits(768, 2112, 131072, split_k=3) == 3
    assert splits(768, 2112, 131072, epilogue=L.EPI_BF16) == 0   # only fp32 outputs are split


def test_exec_model_file_defines_reference_names_and_refuses_cpu():
    """The model file is consumed by exec() (transformer.py:33-34, register.py:11-12, embed.py:164-165): the names it
    must define, and a direct module call on CPU tensors fails loudly (no CPU fallback) instead of NotImplementedError."""
    import numpy as np
    import torch
    import torch.nn as nn
    import torch.nn.functional as F
    from recommendersystem_b200 import lib as L, synthetic
    ns = {"torch": torch, "nn": nn, "F": F, "np": np}
    with open(os.path.join(ROOT, "integration", "transformer.model.py")) as f:
        exec(f.read(), ns)
    assert ns["ALL_MEDIUMS"] == [0, 1] and ns["ALL_METRICS"] == ["watch", "rating"]
    cfg, _ = synthetic.make_config("tiny")
    model = ns["RecommenderModel"](dict(cfg, forward="inference"))
    with pytest.raises(L.GraftError):
        model.item_embedding(torch.arange(0, 4))
    with pytest.raises(L.GraftError):
        model.watch_head((torch.zeros(2, cfg["embed_dim"]), 0))


def test_scratch_scope_owned_by_graph_refuses_to_grow():
    """a captured CUDA graph bakes in workspace pointers; its scope must not hand them back to the allocator.
    A frozen scope raises instead of replacing a buffer; the default scope keeps growing independently."""
    import torch
    scope = L.ScratchScope()
    with scope:
        a = L.scratch(1 << 10, "cpu", slot=7)
        assert L.scratch(1 << 9, "cpu", slot=7) is a          # reused
    scope.frozen = True
    with scope:
        assert L.scratch(1 << 10, "cpu", slot=7) is a
        with pytest.raises(L.GraftError):
            L.scratch(8 << 20, "cpu", slot=7)                  # growth after capture
    b = L.scratch(8 << 20, "cpu", slot=7)                      # the process-wide scope is unaffected
    assert b is not a and scope.bufs[("cpu", 7)] is a
    assert isinstance(b, torch.Tensor)


def test_batch_dtype_normalisation_in_place():
    """the kernels hard-code int32 ids / float64 time / float32 values; the reference indexes with any integer
    dtype, so the drop-in converts instead of reinterpreting (the caller's dict is updated in place, like the mask edits)."""
    import torch
    from recommendersystem_b200.model import _normalise_batch
    d = {"userid": torch.arange(4), "time": torch.ones(4, dtype=torch.float32), "rating": torch.ones(4, dtype=torch.float64),
         "0.watch.position": torch.arange(4), "0.watch.weight": torch.ones(4, dtype=torch.float64), "matchedid": torch.arange(4, dtype=torch.int32)}
    keep = d["matchedid"]
    out = _normalise_batch(d)
    assert out is d and d["matchedid"] is keep                 # already right: untouched
    assert d["userid"].dtype == torch.int32 and d["time"].dtype == torch.float64 and d["rating"].dtype == torch.float32
    assert d["0.watch.position"].dtype == torch.int32 and d["0.watch.weight"].dtype == torch.float32
    with pytest.raises(L.GraftError):
        L.tptr(torch.arange(4), torch.int32, "ids")
