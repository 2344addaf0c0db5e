from densephrases_b200.truecase import TrueCaser  # noqa: F401  (squad_utils.py:1452; the rest of that module is training-side, out of scope)
