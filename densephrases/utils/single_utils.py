from densephrases_b200.runtime import load_encoder  # noqa: F401
