"""Importable stand-in for `faiss`: eval_phrase_retrieval.py:12 and train_query.py:12 import it but never use it, and FAISS
is removed from the hot path (BASELINE.json north_star).  Any attribute access fails loudly."""


def __getattr__(name):
    raise RuntimeError(f'faiss.{name}: FAISS is not part of this build; the search runs on libdph_b200 (densephrases_b200.IvfPqIndex)')
