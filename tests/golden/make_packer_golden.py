"""Writes tests/golden/packer_golden.json: random user histories and what the REFERENCE's
own tokenize / project (notebooks/Finetune/embed.py:39-71, the Python twin of
history_tools.jl:37-75) make of them.  Build container only (needs /root/reference).
    python tests/golden/make_packer_golden.py
"""
import json
import os

import numpy as np

REF = "/root/reference/notebooks/Finetune/embed.py"


def reference_functions():
    src = open(REF).read().split("\n")
    start = next(i for i, l in enumerate(src) if l.startswith("def tokenize("))
    end = next(i for i, l in enumerate(src) if l.startswith("def predict("))
    ns = {}
    exec("\n".join(src[start:end]), ns)  # the two pure functions, verbatim, executed from the reference file
    return ns["tokenize"], ns["project"]


def random_user(rng, n_events):
    items, hist = [], {}
    ts = 1.0e9
    for _ in range(n_events):
        if items and rng.random() < 0.35:  # repeat the previous item: exercises span collapsing
            medium, mid = items[-1]["medium"], items[-1]["matchedid"]
        else:
            medium, mid = int(rng.integers(0, 2)), int(rng.integers(0, 40))
        prev = hist.get((medium, mid))
        status = int(rng.integers(0, 9))
        rating = float(rng.integers(0, 11)) if rng.random() < 0.6 else 0.0
        if prev is not None and rng.random() < 0.25:  # a no-op event: dropped by project
            status, rating = prev
        ts += float(rng.integers(1, 10 ** 6))
        items.append({"medium": medium, "matchedid": mid, "itemid": mid + 1000 * medium, "history_max_ts": ts,
                      "history_tag": "x", "status": status, "rating": rating, "progress": float(np.float32(rng.random())),
                      "history_status": None if prev is None else prev[0],
                      "history_rating": None if prev is None else prev[1]})
        hist[(medium, mid)] = (status, rating)
    g = rng.integers(-1, 3)
    return {"user": {"gender": None if g < 0 else int(g), "source": int(rng.integers(0, 4))}, "items": items}


def main():
    tokenize, project = reference_functions()
    rng = np.random.default_rng(11)
    users = [random_user(rng, int(n)) for n in [0, 1, 2, 5, 17, 40, 120, 300]]
    out = {"users": users, "tokenized": [tokenize(u["items"]) for u in users],
           "projected": [project(tokenize(u["items"])) for u in users]}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "packer_golden.json")
    json.dump(out, open(path, "w"))
    print("wrote", path, os.path.getsize(path) // 1024, "KiB;", [len(p) for p in out["projected"]], "tokens")


if __name__ == "__main__":
    main()
