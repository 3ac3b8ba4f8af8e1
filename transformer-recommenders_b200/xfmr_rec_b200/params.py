"""Constants that matter on the hot path (xfmr_rec/params.py:11-13)."""

METRIC = {"name": "val/retrieval_normalized_dcg", "mode": "max"}
TOP_K = 20
EMBEDDING_DIM = 384  # all-MiniLM-L6-v2 sentence embeddings (params.py:11)
