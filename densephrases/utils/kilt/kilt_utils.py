def store_data(*args, **kwargs):
    raise NotImplementedError('KILT I/O (densephrases/utils/kilt/kilt_utils.py) is benchmark tooling, out of scope (SURVEY.md 2)')
