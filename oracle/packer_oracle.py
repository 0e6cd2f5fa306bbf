"""TEST INFRASTRUCTURE ONLY - literal restatement of the Julia packer's per-user step.

Follows /root/reference/notebooks/Training/history_tools.jl:37-75 (`project!`,
`tokenize!`) and /root/reference/notebooks/Training/transformer.jl:79-142 (`get_data`)
statement by statement, on the same msgpack-shaped dicts (Julia `nothing` = None).
Julia is not installed in the build image, so this half of the path cannot be pinned by
running the reference; it is pinned instead against the reference's own Python
restatement of tokenize/project (/root/reference/notebooks/Finetune/embed.py:39-71),
executed by tests/golden/make_packer_golden.py -> tests/golden/packer_golden.json.
Only tests/ may import this file.
"""
import numpy as np

PLANNED_STATUS = 5                         # transformer.jl:18
METRICS = ("watch", "rating", "status")    # transformer.jl:17


def tokenize(items):
    """history_tools.jl:48-75: consecutive events on one (medium, matchedid) collapse to one
    token: first event's fields, last event's status / rating / progress."""
    out, span, last_mid = [], [], None

    def span_to_token(sp):
        tok = dict(sp[0])
        for k in ("status", "rating", "progress"):
            tok[k] = sp[-1][k]
        return tok

    for x in items:
        mid = (x["medium"], x["matchedid"])
        if mid == last_mid:
            span.append(x)
        else:
            if span:
                out.append(span_to_token(span))
            span, last_mid = [x], mid
    if span:
        out.append(span_to_token(span))
    return out


def project(items):
    """history_tools.jl:37-46: drop events that change neither status nor rating."""
    return [x for x in items if not (x["history_status"] == x["status"] and x["history_rating"] == x["rating"])]


def get_data(data, userid, num_items0):
    """transformer.jl:79-142.  `num_items0` = num_items(0), the manga catalog size (:32-35)."""
    items = project(tokenize(data["items"]))
    u = data["user"]
    n = len(items)
    d = {
        "userid": np.empty(n, np.int32), "token_mask_ids": np.zeros(n, np.int32), "time": np.empty(n, np.float64),
        "gender": np.empty(n, np.int32), "source": np.empty(n, np.int32), "matchedid": np.empty(n, np.int32),
        "status": np.empty(n, np.int32), "rating": np.empty(n, np.float32), "progress": np.empty(n, np.float32),
    }
    for m in (0, 1):
        for metric in METRICS:
            d[f"{m}.{metric}.label"] = np.zeros(n, np.float32)
            d[f"{m}.{metric}.weight"] = np.zeros(n, np.float32)
            d[f"{m}.{metric}.position"] = np.zeros(n, np.int32)
    for i, x in enumerate(items):
        m = x["medium"]
        d["userid"][i] = userid
        d["time"][i] = x["history_max_ts"]
        d["gender"][i] = 0 if u["gender"] is None else u["gender"] + 1
        d["source"][i] = u["source"]
        d["matchedid"][i] = x["matchedid"] + (num_items0 if m == 1 else 0)
        d["status"][i] = x["status"]
        d["rating"][i] = x["rating"]
        d["progress"][i] = x["progress"]
        hs, hr = x["history_status"], x["history_rating"]
        inferred_watch = x["status"] == 0 and hs is None
        new_watch = x["status"] > PLANNED_STATUS and (hs is None or 0 < hs <= PLANNED_STATUS)
        if inferred_watch or new_watch:
            d[f"{m}.watch.label"][i] = 1
            d[f"{m}.watch.weight"][i] = 1
            d[f"{m}.watch.position"][i] = x["matchedid"]
        if x["rating"] > 0 and x["rating"] != hr:
            d["token_mask_ids"][i] = 1
            d[f"{m}.rating.label"][i] = x["rating"]
            d[f"{m}.rating.weight"][i] = 1
            d[f"{m}.rating.position"][i] = x["matchedid"]
        if x["status"] > 0 and x["status"] != hs:
            d[f"{m}.status.label"][i] = x["status"]
            d[f"{m}.status.weight"][i] = 1
            d[f"{m}.status.position"][i] = x["matchedid"]
    return d


def get_data_finetune(data, userid, num_items0, max_seq_len=1024, num_test_items=1):
    """/root/reference/notebooks/Finetune/transformer.jl:52-133 statement by statement: the user's history (tokenized and
    projected, truncated to the last max_seq_len - num_test_items tokens) followed by the held-out test items, in a
    fixed-length row of max_seq_len positions (zero padded); only test items carry targets; each (medium, metric)
    weight vector is normalised to sum 1 (:124-131)."""
    items = project(tokenize(data["items"]))
    assert len(data["test_items"]) <= num_test_items
    N = max_seq_len
    d = {"userid": np.zeros(N, np.int32), "token_mask_ids": np.zeros(N, np.int32), "time": np.zeros(N, np.float64),
         "gender": np.zeros(N, np.int32), "source": np.zeros(N, np.int32), "matchedid": np.zeros(N, np.int32),
         "status": np.zeros(N, np.int32), "rating": np.zeros(N, np.float32), "progress": np.zeros(N, np.float32)}
    for m in (0, 1):
        for metric in METRICS:
            d[f"{m}.{metric}.label"] = np.zeros(N, np.float32)
            d[f"{m}.{metric}.weight"] = np.zeros(N, np.float32)
            d[f"{m}.{metric}.position"] = np.zeros(N, np.int32)
    if len(items) > max_seq_len - num_test_items:
        items = items[len(items) - (max_seq_len - num_test_items):]
    u = data["user"]
    i = 0
    for source, istest in ((items, False), (data["test_items"], True)):
        for x in source:
            m = x["medium"]
            d["userid"][i] = userid
            d["time"][i] = x["history_max_ts"]
            d["gender"][i] = 0 if u["gender"] is None else u["gender"] + 1
            d["source"][i] = u["source"]
            d["matchedid"][i] = x["matchedid"] + (num_items0 if m == 1 else 0)
            d["status"][i] = x["status"]
            d["rating"][i] = x["rating"]
            d["progress"][i] = x["progress"]
            if istest:
                hs, hr = x["history_status"], x["history_rating"]
                inferred_watch = x["status"] == 0 and hs is None
                new_watch = x["status"] > PLANNED_STATUS and (hs is None or 0 < hs <= PLANNED_STATUS)
                if inferred_watch or new_watch:
                    d[f"{m}.watch.label"][i] = 1
                    d[f"{m}.watch.weight"][i] = 1
                    d[f"{m}.watch.position"][i] = x["matchedid"]
                if x["rating"] > 0 and x["rating"] != hr:
                    d["token_mask_ids"][i] = 1
                    d[f"{m}.rating.label"][i] = x["rating"]
                    d[f"{m}.rating.weight"][i] = 1
                    d[f"{m}.rating.position"][i] = x["matchedid"]
                if x["status"] > 0 and x["status"] != hs:
                    d[f"{m}.status.label"][i] = x["status"]
                    d[f"{m}.status.weight"][i] = 1
                    d[f"{m}.status.position"][i] = x["matchedid"]
            i += 1
    for m in (0, 1):
        for metric in METRICS:
            wsum = d[f"{m}.{metric}.weight"].sum(dtype=np.float32)
            if wsum > 0:
                d[f"{m}.{metric}.weight"] = d[f"{m}.{metric}.weight"] / wsum
    return d
