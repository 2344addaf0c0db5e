from densephrases_b200.runtime import backward_compat, load_encoder  # noqa: F401
