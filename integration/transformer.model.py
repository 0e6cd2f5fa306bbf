# Drop-in for /root/reference/notebooks/Training/transformer.model.py.
#
# The reference's callers paste the model file into their own namespace:
#     with open("transformer.model.py") as f: exec(f.read())
# (notebooks/Training/transformer.py:33-34, notebooks/Finetune/register.py:11-12,
# notebooks/Finetune/embed.py:164-165).  Copy this file over the reference's
# transformer.model.py (or next to the trainer) and the trainer runs unchanged:
# the names below are the ones it uses (transformer.py:209-212, 254, 391, 403,
# 483, 647), every module and parameter keeps its reference name, so
# checkpoints written by either implementation load in the other.
#
# The step itself runs on libgraft.so (sm_100a); there is no CPU fallback.
from recommendersystem_b200.model import (  # noqa: F401
    ALL_MEDIUMS,
    ALL_METRICS,
    RecommenderModel,
    init_weights,
)
