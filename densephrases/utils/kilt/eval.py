def evaluate(*args, **kwargs):
    raise NotImplementedError('KILT scoring (densephrases/utils/kilt/eval.py) is benchmark tooling, out of scope (SURVEY.md 2)')
