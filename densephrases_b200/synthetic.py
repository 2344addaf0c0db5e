"""Synthetic corpus / phrase-index builders (no dataset or checkpoint is reachable offline).

`make_corpus` produces the three artefacts MIPS consumes besides the vector index: idx2id (label -> doc, token),
per-document metadata (word2char_start/end, f2o_start, context, title) and the doc-major label numbering of
build_phrase_index.py:145-150.  `make_phrase_index_arrays` assigns every token vector a random inverted list and a
random PQ96 code and returns list-major arrays for IvfPqIndex.from_arrays."""
import numpy as np

_WORDS = ("alpha beta gamma delta epsilon zeta eta theta iota kappa lambda mu nu xi omicron pi rho sigma tau upsilon phi chi psi omega "
          "river mountain city treaty album battle empire railway species theory island museum council engine harbor senate").split()


def make_corpus(n_docs=20, seed=0, min_words=40, max_words=160):
    rng = np.random.default_rng(seed)
    doc_groups, docs, words = {}, [], []
    for d in range(n_docs):
        n_words = int(rng.integers(min_words, max_words))
        toks, starts, ends, pos = [], [], [], 0
        pieces = []
        for w in range(n_words):
            word = _WORDS[int(rng.integers(len(_WORDS)))]
            if rng.random() < 0.08:
                word = word.capitalize()
            r = rng.random()
            sep = ' [PAR] ' if (r < 0.03 and w > 0) else ('. ' if r < 0.12 and w > 0 else (' ' if w > 0 else ''))
            if sep == '. ':
                pieces.append('.')
                pos += 1
                sep = ' '
            pieces.append(sep)
            pos += len(sep)
            starts.append(pos)
            pieces.append(word)
            pos += len(word)
            ends.append(pos)
            toks.append(word)
        context = ''.join(pieces)
        # features (word pieces): every word yields 1 or 2 features; f2o_start maps feature -> original word
        reps = rng.integers(1, 3, n_words)
        f2o = np.repeat(np.arange(n_words), reps).astype(np.int32)
        doc_groups[str(d)] = {'word2char_start': np.array(starts, dtype=np.int32), 'word2char_end': np.array(ends, dtype=np.int32),
                              'f2o_start': f2o, 'context': context, 'title': f'Doc {d}'}
        docs.append(np.full(len(f2o), d, dtype=np.int32))
        words.append(np.arange(len(f2o), dtype=np.int32))
    idx_f = {'0': {'doc': np.concatenate(docs), 'word': np.concatenate(words)}}      # one offset group (build_phrase_index.py:268-276)
    return doc_groups, idx_f, int(sum(len(x) for x in docs))


def make_phrase_index_arrays(ntotal, nlist, seed=0):
    """Random list assignment + random codes for `ntotal` doc-major labels -> (list_len, codes, ids) list-major."""
    rng = np.random.default_rng(seed + 17)
    assign = rng.integers(0, nlist, ntotal)
    order = np.argsort(assign, kind='stable')
    list_len = np.bincount(assign, minlength=nlist).astype(np.int64)
    codes = rng.integers(0, 256, (ntotal, 96), dtype=np.uint8)
    return list_len, codes[order], order.astype(np.int64)          # ids = doc-major label of each list-major row
