"""Synthetic phrase dumps at full scale WITHOUT the bytes on disk (BASELINE.json configs[4]).

A real DensePhrases dump is `dump_dir/<index_name>/index.faiss` + `idx2id.hdf5` + `dump_dir/meta_compressed.pkl`
(open_utils.py:28-31, index.py:24-76).  A multi_wiki-scale synthetic rebuild would be 55 GB of codes, 4.6 GB of idx2id and
tens of GB of metadata; none of it carries information (codes are uniform bytes from a seeded counter-based generator), so a
synthetic dump is three tiny JSON *spec* files that `MIPS` understands:

    dump_dir/<index_name>/index.dph.json     {"synthetic_index": {N, nlist, seed, opq_seed}}   -> generated on the GPU(s)
    dump_dir/<index_name>/idx2id.dph.json    {"synthetic_idx2id": {ntotal, tokens_per_doc}}     -> arithmetic lookups
    dump_dir/meta_dph.json                   {"synthetic_meta": {tokens_per_doc, seed}}         -> documents made on demand

Labels follow build_phrase_index.py:145-150 (label = position of the token vector in dump order = doc-major), documents have
a fixed number of token vectors, and every document's text / word offsets / f2o_start are a pure function of (seed, doc_idx),
so every rank of a sharded job sees the same corpus.  `write_synthetic_dump` lays the three files out the way
`load_phrase_index` expects them."""
import json
import os

import numpy as np

_WORDS = ("alpha beta gamma delta epsilon zeta eta theta iota kappa lambda mu nu xi omicron pi rho sigma tau upsilon phi chi psi omega "
          "river mountain city treaty album battle empire railway species theory island museum council engine harbor senate").split()


class _Affine(object):
    """`idx_f[offset]['doc'][rows]` / `['word'][rows]` of a corpus with a fixed number of token vectors per document."""

    def __init__(self, ntotal, per_doc, kind):
        self.ntotal, self.per_doc, self.kind = int(ntotal), int(per_doc), kind

    def __len__(self):
        return self.ntotal

    def __getitem__(self, rows):
        rows = np.asarray(rows, dtype=np.int64)
        return rows // self.per_doc if self.kind == 'doc' else rows % self.per_doc


def synthetic_idx2id(ntotal, tokens_per_doc):
    """{offset_key: {'doc','word'}} like index.py:78-88 with ONE offset group (labels < 1e9 for PQ indexes, index.py:33)."""
    return {'0': {'doc': _Affine(ntotal, tokens_per_doc, 'doc'), 'word': _Affine(ntotal, tokens_per_doc, 'word')}}


class LazyDocs(object):
    """doc_groups[str(doc_idx)] -> the record MIPS.decompress_meta reads (index.py:106-122), generated on demand and cached."""

    def __init__(self, tokens_per_doc, seed=0, cache=4096):
        self.per_doc, self.seed, self.cache_cap = int(tokens_per_doc), int(seed), cache
        self._cache = {}

    def __contains__(self, key):
        return int(key) >= 0

    def __getitem__(self, key):
        d = int(key)
        rec = self._cache.get(d)
        if rec is None:
            rec = self._make(d)
            if len(self._cache) >= self.cache_cap:
                self._cache.pop(next(iter(self._cache)))
            self._cache[d] = rec
        return rec

    def _make(self, d):
        rng = np.random.default_rng((self.seed << 32) ^ d)
        n = self.per_doc
        widx = rng.integers(0, len(_WORDS), n)
        sep = rng.random(n)
        pieces, starts, ends, pos = [], np.empty(n, np.int32), np.empty(n, np.int32), 0
        for w in range(n):
            if w:
                s = ' [PAR] ' if sep[w] < 0.03 else ('. ' if sep[w] < 0.12 else ' ')
                pieces.append(s)
                pos += len(s)
            word = _WORDS[widx[w]]
            starts[w] = pos
            pieces.append(word)
            pos += len(word)
            ends[w] = pos
        return {'word2char_start': starts, 'word2char_end': ends, 'f2o_start': np.arange(n, dtype=np.int32),
                'context': ''.join(pieces), 'title': f'Doc {d}'}


def write_synthetic_dump(dump_dir, index_name, N, nlist, tokens_per_doc=128, seed=1234, phrase_dir='phrase'):
    """Create the directory layout of `load_phrase_index` (open_utils.py:28-31) with spec files only."""
    index_dir = os.path.join(dump_dir, index_name)
    os.makedirs(index_dir, exist_ok=True)
    os.makedirs(os.path.join(dump_dir, phrase_dir), exist_ok=True)
    ntotal = (int(N) // tokens_per_doc) * tokens_per_doc
    json.dump({'synthetic_index': {'N': ntotal, 'nlist': int(nlist), 'seed': int(seed), 'opq_seed': int(seed)}}, open(os.path.join(index_dir, 'index.dph.json'), 'w'))
    json.dump({'synthetic_idx2id': {'ntotal': ntotal, 'tokens_per_doc': tokens_per_doc}}, open(os.path.join(index_dir, 'idx2id.dph.json'), 'w'))
    json.dump({'synthetic_meta': {'tokens_per_doc': tokens_per_doc, 'seed': int(seed)}}, open(os.path.join(dump_dir, 'meta_dph.json'), 'w'))
    return ntotal


def write_synthetic_questions(path, n, seed=0):
    """A QA file in the format load_qa_pairs reads (open_utils.py:104-160): synthetic questions over the corpus vocabulary."""
    rng = np.random.default_rng(seed)
    data = []
    for i in range(n):
        q = ' '.join(_WORDS[j] for j in rng.integers(0, len(_WORDS), int(rng.integers(4, 12)))) + '?'
        data.append({'id': f'syn-{i}', 'question': q, 'answers': [_WORDS[int(rng.integers(len(_WORDS)))]]})
    json.dump({'data': data}, open(path, 'w'))
    return path


def uniform_list_lengths(N, nlist):
    base, rem = divmod(int(N), int(nlist))
    lens = np.full(int(nlist), base, dtype=np.int64)
    lens[:rem] += 1
    return lens
