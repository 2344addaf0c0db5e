"""Read access to the reference's phrase dump -- HDF5 files with one group per document holding the int8 token vectors (`start`),
the word/char maps (`word2char_start`, `word2char_end`, `f2o_start`) and the attributes `context` / `title`
(written by densephrases/utils/embed_utils.py:233-246, read by index.py:90-104,143-156,246-273).

`MIPS.search_phrase` falls back to it when there is no in-RAM metadata (`meta_compressed.pkl`): functional, not accelerated
(SURVEY.md 8a "explicitly outside").  Backends, in order: an in-memory mapping (tests, synthetic dumps), h5py when it is importable,
else this repo's native HDF5 subset reader (artifacts.py)."""
import os

import numpy as np

FIELDS = ('start', 'word2char_start', 'word2char_end', 'f2o_start')


class DictPhraseDump(object):
    """{str(doc_idx) or doc_idx: {'start': int8 [T,768], 'word2char_start', 'word2char_end', 'f2o_start', 'context', 'title'}}."""

    def __init__(self, records):
        self.records = records

    def get(self, doc_idx):
        rec = self.records.get(str(doc_idx), None)
        if rec is None:
            rec = self.records.get(int(doc_idx), None)
        if rec is None:
            raise ValueError('%d not found in dump list' % int(doc_idx))          # index.py:151,155
        return rec


class PhraseDump(object):
    """The dump files of `phrase_dump_dir` (a directory of *.hdf5 files, or one file), opened on first use.  File selection follows
    index.py:90-100,143-156: with several files named `<lo>-<hi>.hdf5` a document lives in the file whose range (in thousands)
    covers it, else in the last file."""

    def __init__(self, phrase_dump_dir):
        self.phrase_dump_dir = phrase_dump_dir
        self.files = None
        self.ranges = None

    def _open(self):
        d = self.phrase_dump_dir
        if d is None or not os.path.exists(d):
            raise NotImplementedError(f'phrase dump {d!r} does not exist: there is no in-RAM metadata and no dump to read token '
                                      'vectors from (index.py:246-273)')
        paths = sorted(os.path.join(d, n) for n in os.listdir(d) if 'hdf5' in n) if os.path.isdir(d) else [d]
        if not paths:
            raise NotImplementedError(f'no *.hdf5 phrase dump under {d!r} (index.py:90-100)')
        names = [os.path.splitext(os.path.basename(p))[0] for p in paths]
        if '-' in names[0] and 'dev' not in names[0]:
            self.ranges = [list(map(int, n.split('-'))) for n in names]
        self.files = [_open_file(p) for p in paths]

    def _group(self, doc_idx):
        if self.files is None:
            self._open()
        key = str(doc_idx)
        if len(self.files) == 1:
            return self.files[0].group(key)
        if self.ranges is not None:
            for (lo, hi), f in zip(self.ranges, self.files):
                if lo * 1000 <= int(doc_idx) < hi * 1000:
                    if not f.has(key):
                        raise ValueError('%d not found in dump list' % int(doc_idx))
                    return f.group(key)
        if not self.files[-1].has(key):
            raise ValueError('%d not found in dump list' % int(doc_idx))
        return self.files[-1].group(key)

    def get(self, doc_idx):
        return self._group(doc_idx)


def _open_file(path):
    try:
        import h5py
        if hasattr(h5py, 'File'):
            return _H5pyFile(h5py.File(path, 'r'))
    except ImportError:
        pass
    return _NativeFile(path)


class _H5pyFile(object):
    def __init__(self, f):
        self.f = f

    def has(self, key):
        return key in self.f

    def group(self, key):
        g = self.f[key]
        rec = {name: g[name] for name in FIELDS}          # `start` stays a lazy dataset (sliced per hit), the maps are small
        for name in FIELDS[1:]:
            rec[name] = rec[name][:]
        rec['context'], rec['title'] = _text(g.attrs['context']), _text(g.attrs['title'])
        return rec


class _NativeFile(object):
    """artifacts.py's HDF5 subset reader: old-style groups, contiguous / chunked datasets, string attributes in the object header."""

    def __init__(self, path):
        from . import artifacts
        self.h = artifacts.open_hdf5(path)
        self.root = dict(self.h.children(self.h.root_addr))

    def has(self, key):
        return key in self.root

    def group(self, key):
        members = dict(self.h.children(self.root[key]))
        rec = {name: self.h.dataset(members[name]) for name in FIELDS}
        attrs = self.h.attributes(self.root[key])
        rec['context'], rec['title'] = _text(attrs['context']), _text(attrs['title'])
        return rec


def _text(v):
    if isinstance(v, bytes):
        return v.decode('utf-8')
    if isinstance(v, np.ndarray):
        return _text(v.item() if v.shape == () else v.tolist()[0])
    return str(v)
