"""MIPS API mirror (densephrases_b200/mips.py) vs the item-by-item restatement of the reference's phrase stage
(oracle/mips_ref.py).  CPU: host logic over an adapter around the oracle index (the CUDA index cannot run here);
GPU: the same comparison end to end through libdph_b200."""
import numpy as np
import pytest

from tests.helpers import opq_matrix

NLIST, SEED = 24, 3


class OracleIndexAdapter:
    """Test-only: gives the oracle RefIndex the IvfPqIndex surface MIPS uses (search / reconstruct_batch / ...)."""

    def __init__(self, ref):
        self.ref, self.nprobe, self.d, self.ntotal = ref, 256, ref.d, ref.ntotal

    def search(self, x, k):
        return self.ref.search(x, k, self.nprobe)

    def reconstruct_batch(self, ids):
        return self.ref.reconstruct(np.asarray(ids, dtype=np.int64))

    def opq_matrix(self):
        return self.ref.A


def build(oracle, n_docs=12):
    from densephrases_b200.synthetic import make_corpus, make_phrase_index_arrays
    doc_groups, idx_f, ntotal = make_corpus(n_docs, SEED)
    list_len, codes, ids = make_phrase_index_arrays(ntotal, NLIST, SEED)
    A, pq, Cm = opq_matrix(SEED), oracle.gen_pq(SEED), oracle.gen_centroids(SEED, 0, NLIST)
    ref = oracle.RefIndex(A, pq, list_len, centroids=Cm, codes=codes, ids=ids)
    rng = np.random.default_rng(9)
    pick = rng.integers(0, ntotal, (6, 2))
    vs, _ = ref.reconstruct(pick[:, 0])
    ve, _ = ref.reconstruct(np.minimum(pick[:, 0] + rng.integers(0, 4, 6), ntotal - 1))
    query = np.concatenate([vs @ A, ve @ A], 1).astype(np.float64) + 0.05 * rng.standard_normal((6, 1536))
    return doc_groups, idx_f, (A, pq, Cm, list_len, codes, ids), ref, query


def dump_records(doc_groups, seed=21):
    """The synthetic corpus in the phrase-dump form (one record per document: int8 token vectors `start`, the maps, context / title),
    shared by the golden generator (served to the UNMODIFIED reference through a stand-in h5py) and the tests."""
    rng = np.random.default_rng(seed)
    recs = {}
    for k, g in doc_groups.items():
        T = len(g['f2o_start'])
        recs[str(k)] = {'start': rng.integers(-128, 128, (T, 768), dtype=np.int8), 'word2char_start': np.asarray(g['word2char_start']),
                        'word2char_end': np.asarray(g['word2char_end']), 'f2o_start': np.asarray(g['f2o_start']),
                        'context': g['context'], 'title': g['title']}
    return recs


def compare(outs, refs, vec_tol=None):
    assert len(outs) == len(refs)
    for got, want in zip(outs, refs):
        assert len(got) == len(want)
        for g, w in zip(got, want):
            for key in ('context', 'title', 'doc_idx', 'start_pos', 'end_pos', 'start_idx', 'end_idx', 'answer'):
                assert g[key] == w[key], (key, g[key], w[key])
            assert abs(g['score'] - w['score']) <= 1e-3 * max(1.0, abs(w['score']))        # fp32 matmul order differs (torch vs numpy)
            if vec_tol is not None:
                assert np.abs(np.asarray(g['start_vec']) - w['start_vec']).max() < vec_tol
                assert np.abs(np.asarray(g['end_vec']) - w['end_vec']).max() < vec_tol


@pytest.mark.parametrize("aggregate,agg", [(False, 'opt1'), (True, 'opt1'), (True, 'opt2'), (True, 'opt3'), (True, 'opt4')])
def test_mips_host_logic_matches_reference_restatement_cpu(oracle, aggregate, agg):
    from densephrases_b200.mips import MIPS, normalize_answer
    from oracle.mips_ref import ref_search
    doc_groups, idx_f, _, ref, query = build(oracle)
    mips = MIPS.from_components(OracleIndexAdapter(ref), idx_f, doc_groups, cuda=False)
    mips.index.nprobe = 8
    outs = mips.search(query, q_texts=['q'] * len(query), top_k=5, aggregate=aggregate, agg_strat=agg, return_idxs=True)
    refs = ref_search(ref, idx_f, doc_groups, query, top_k=5, nprobe=8, aggregate=aggregate, agg_strat=agg, return_idxs=True,
                      normalize_answer=normalize_answer)
    compare(outs, refs, vec_tol=1e-3)
    assert len(mips.num_docs_list) == 1


def test_fused_window_score_branch_on_cpu(oracle):
    """The branch MIPS takes on the GPU (index.window_scores: reconstruct + un-rotate + dot fused in one call, return_idxs=False),
    driven on the CPU by an adapter that computes the same window scores from the oracle: same results as the reference restatement."""
    from densephrases_b200.mips import MIPS, normalize_answer
    from oracle.mips_ref import ref_search

    class Fused(OracleIndexAdapter):
        calls = 0

        def window_scores(self, q, first_ids, L):
            Fused.calls += 1
            xq = self.ref.rotate(np.ascontiguousarray(q, dtype=np.float32))          # <q, A^T v> == <A q, v>
            out = np.zeros((len(first_ids), L), dtype=np.float32)
            for j in range(L):
                v, _ = self.ref.reconstruct(np.asarray(first_ids, dtype=np.int64) + j)     # missing label -> zero row
                out[:, j] = (xq * v).sum(1)
            return out
    doc_groups, idx_f, _, ref, query = build(oracle)
    mips = MIPS.from_components(Fused(ref), idx_f, doc_groups, cuda=False)
    mips.index.nprobe = 8
    outs = mips.search(query, q_texts=['q'] * len(query), top_k=5, aggregate=True, agg_strat='opt1', return_idxs=False)
    assert Fused.calls == 2                                                          # one fused call per direction
    refs = ref_search(ref, idx_f, doc_groups, query, top_k=5, nprobe=8, aggregate=True, agg_strat='opt1', normalize_answer=normalize_answer)
    compare(outs, refs)


def test_get_idxs_clips_out_of_range(oracle):
    from densephrases_b200.mips import MIPS
    doc_groups, idx_f, _, ref, _ = build(oracle)
    mips = MIPS.from_components(OracleIndexAdapter(ref), idx_f, doc_groups, cuda=False)
    doc, word = mips.get_idxs(np.array([[-1, 0, ref.ntotal + 5]]))
    assert doc[0, 0] == idx_f['0']['doc'][0] and doc[0, 2] == idx_f['0']['doc'][ref.ntotal - 1] and word[0, 1] == 0


def test_sentence_crop_and_zlib_metadata(oracle):
    import zlib
    from densephrases_b200.mips import MIPS
    doc_groups, idx_f, _, ref, query = build(oracle)
    packed = {}
    for k, g in doc_groups.items():     # the reference keeps compressed blobs + dtypes (compress_metadata.py:32-53)
        packed[k] = {'word2char_start': zlib.compress(g['word2char_start'].tobytes()), 'word2char_end': zlib.compress(g['word2char_end'].tobytes()),
                     'f2o_start': zlib.compress(g['f2o_start'].tobytes()), 'context': zlib.compress(g['context'].encode()), 'title': g['title'],
                     'dtypes': {'word2char_start': g['word2char_start'].dtype, 'word2char_end': g['word2char_end'].dtype, 'f2o_start': g['f2o_start'].dtype}}
    a = MIPS.from_components(OracleIndexAdapter(ref), idx_f, doc_groups, cuda=False)
    b = MIPS.from_components(OracleIndexAdapter(ref), idx_f, packed, cuda=False)
    a.index.nprobe = b.index.nprobe = 8
    ra = a.search(query, q_texts=['q'] * 6, top_k=4, return_sent=True)
    rb = b.search(query, q_texts=['q'] * 6, top_k=4, return_sent=True)
    for x, y in zip(ra, rb):
        assert [r['context'] for r in x] == [r['context'] for r in y]
        for r in x:
            assert r['context'][r['start_pos']:r['end_pos']] == r['answer']


@pytest.mark.gpu
def test_mips_end_to_end_on_gpu(oracle):
    from densephrases_b200 import IvfPqIndex
    from densephrases_b200.mips import MIPS, normalize_answer
    from oracle.mips_ref import ref_search
    doc_groups, idx_f, (A, pq, Cm, list_len, codes, ids), ref, query = build(oracle, n_docs=40)
    index = IvfPqIndex.from_arrays(A, Cm, pq, list_len, codes, ids)
    mips = MIPS.from_components(index, idx_f, doc_groups, cuda=True)
    assert mips.index.nprobe == 256                                  # fixed at load (index.py:53,62)
    outs = mips.search(query, q_texts=['q'] * len(query), top_k=10, aggregate=True, agg_strat='opt1', return_idxs=True, nprobe=3)
    refs = ref_search(ref, idx_f, doc_groups, query, top_k=10, nprobe=256, aggregate=True, agg_strat='opt1', return_idxs=True,
                      normalize_answer=normalize_answer)
    compare(outs, refs, vec_tol=1e-3)


@pytest.mark.gpu
def test_fused_window_scores_match_reconstruct_path(oracle):
    """dph_index_window_scores (fused reconstruct + un-rotate + dot) == reconstruct_batch + R matmul + dot; missing labels -> 0."""
    from densephrases_b200 import IvfPqIndex
    from densephrases_b200.mips import MIPS, normalize_answer
    from oracle.mips_ref import ref_search
    doc_groups, idx_f, (A, pq, Cm, list_len, codes, ids), ref, query = build(oracle, n_docs=25)
    index = IvfPqIndex.from_arrays(A, Cm, pq, list_len, codes, ids)
    rng = np.random.default_rng(0)
    first = np.concatenate([rng.integers(0, ref.ntotal, 40), [-3, ref.ntotal - 2, ref.ntotal + 7]]).astype(np.int64)
    q = rng.standard_normal((len(first), 768)).astype(np.float32)
    got = index.window_scores(q, first, 10)
    lab = (first[:, None] + np.arange(10)[None, :]).ravel()
    vec, found = ref.reconstruct(lab)
    want = ((vec.astype(np.float64) @ A.astype(np.float64)).reshape(len(first), 10, 768) * q[:, None, :].astype(np.float64)).sum(2)
    assert np.abs(got - want).max() < 1e-3 * max(1.0, np.abs(want).max())
    assert (got.ravel()[found == 0] == 0).all()
    mips = MIPS.from_components(index, idx_f, doc_groups, cuda=True)
    outs = mips.search(query, q_texts=['q'] * len(query), top_k=10, aggregate=True, agg_strat='opt1', return_idxs=False)      # fused path
    refs = ref_search(ref, idx_f, doc_groups, query, top_k=10, nprobe=256, aggregate=True, agg_strat='opt1', normalize_answer=normalize_answer)
    compare(outs, refs)


@pytest.mark.parametrize("which", ["restatement", "mips"])
def test_phrase_stage_matches_reference_golden(oracle, which):
    """tests/golden/mips_search.json holds what the UNMODIFIED reference `MIPS.search` (index.py:124-141,189-482) returned for
    this corpus / index / query batch (generator: tests/golden/make_mips_golden.py, run in the build container).  Both the
    item-by-item restatement used as the GPU-box oracle (oracle/mips_ref.py) and the batched MIPS mirror reproduce it."""
    import json
    import os
    from densephrases_b200.mips import MIPS, normalize_answer
    from oracle.mips_ref import ref_search
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "mips_search.json")))
    doc_groups, idx_f, _, ref, query = build(oracle)
    if which == "mips":                                       # out-of-range labels: clipped like index.py:128-133
        m = MIPS.from_components(OracleIndexAdapter(ref), idx_f, doc_groups, cuda=False)
        doc, word = m.get_idxs(np.array(gold["get_idxs"]["I"], dtype=np.int64))
        assert doc.tolist() == gold["get_idxs"]["doc"] and word.tolist() == gold["get_idxs"]["word"]
    for cfg in gold["configs"]:
        kw = dict(top_k=gold["top_k"], aggregate=cfg["aggregate"], agg_strat=cfg["agg_strat"], return_idxs=True)
        if which == "restatement":
            outs = ref_search(ref, idx_f, doc_groups, query, nprobe=gold["nprobe"], normalize_answer=normalize_answer, **kw)
        else:
            mips = MIPS.from_components(OracleIndexAdapter(ref), idx_f, doc_groups, cuda=False)
            mips.index.nprobe = gold["nprobe"]
            outs = mips.search(query, q_texts=["q"] * len(query), **kw)
        assert len(outs) == len(cfg["results"])
        for got, want in zip(outs, cfg["results"]):
            assert len(got) == len(want), (cfg["agg_strat"], len(got), len(want))
            for g, w in zip(got, want):
                for key in ("context", "title", "doc_idx", "start_pos", "end_pos", "start_idx", "end_idx", "answer"):
                    assert g[key] == w[key], (cfg["agg_strat"], key, g[key], w[key])
                assert abs(g["score"] - w["score"]) <= 1e-3 * max(1.0, abs(w["score"]))
                assert abs(float(np.sum(g["start_vec"])) - w["start_vec_sum"]) < 5e-2 and abs(float(np.sum(g["end_vec"])) - w["end_vec_sum"]) < 5e-2


@pytest.mark.parametrize("source", ["mapping", "native_hdf5_file"])
def test_phrase_dump_branch_matches_reference_golden(oracle, source, tmp_path):
    """Without in-RAM metadata the reference reads token vectors (int8 `start` rows, dequantised x / 20 - 2) and metadata from the phrase
    dump (index.py:246-273).  tests/golden/mips_search_hdf5.json holds what the UNMODIFIED reference `MIPS.search` returned on that
    branch (generator: make_mips_golden.py, a stand-in h5py serving these records); our MIPS reproduces it from an in-memory mapping
    and from an HDF5 file written and read back by this repo's native subset writer / reader (groups, int8 datasets, string attributes)."""
    import json
    import os
    from densephrases_b200 import artifacts
    from densephrases_b200.mips import MIPS
    from densephrases_b200.phrase_dump import DictPhraseDump, PhraseDump
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "mips_search_hdf5.json")))
    doc_groups, idx_f, _, ref, query = build(oracle)
    recs = dump_records(doc_groups)
    if source == "mapping":
        dump = DictPhraseDump(recs)
    else:
        path = str(tmp_path / "phrase" / "0-1.hdf5")
        os.makedirs(os.path.dirname(path))
        artifacts.write_hdf5(path, {k: {f: r[f] for f in ("start", "word2char_start", "word2char_end", "f2o_start")} for k, r in recs.items()},
                             attrs={k: {"context": r["context"], "title": r["title"]} for k, r in recs.items()})
        dump = PhraseDump(os.path.dirname(path))
    for cfg in gold["configs"]:
        mips = MIPS.from_components(OracleIndexAdapter(ref), idx_f, None, cuda=False, phrase_dump=dump)
        mips.index.nprobe = gold["nprobe"]
        outs = mips.search(query, q_texts=["q"] * len(query), top_k=gold["top_k"], aggregate=cfg["aggregate"], agg_strat=cfg["agg_strat"],
                           return_idxs=cfg["return_idxs"])
        assert len(outs) == len(cfg["results"])
        for got, want in zip(outs, cfg["results"]):
            assert len(got) == len(want), (cfg, len(got), len(want))
            for g, w in zip(got, want):
                for key in ("context", "title", "doc_idx", "start_pos", "end_pos", "start_idx", "end_idx", "answer"):
                    assert g[key] == w[key], (cfg["agg_strat"], key, g[key], w[key])
                assert abs(g["score"] - w["score"]) <= 1e-3 * max(1.0, abs(w["score"]))
                if cfg["return_idxs"]:
                    assert abs(float(np.sum(g["start_vec"])) - w["start_vec_sum"]) <= 2e-3 * max(1.0, abs(w["start_vec_sum"]))
                    assert abs(float(np.sum(g["end_vec"])) - w["end_vec_sum"]) <= 2e-3 * max(1.0, abs(w["end_vec_sum"]))
    with pytest.raises(NotImplementedError):                   # neither in-RAM metadata nor a dump
        MIPS.from_components(OracleIndexAdapter(ref), idx_f, None, cuda=False).search(query, q_texts=["q"] * len(query), top_k=3)


@pytest.mark.gpu
def test_mips_loads_a_synthetic_dump_and_equals_the_oracle(oracle, tmp_path):
    """The spec-file dump of densephrases_b200/synthetic_dump.py (C5 at full scale) loaded through the reference's own path
    convention (load_phrase_index -> MIPS(phrase_dump_dir, index_path, idx2id_path)): the GPU-generated index equals the oracle's
    synthetic index (same seed), idx2id is arithmetic, documents are generated on demand; results == the reference restatement."""
    import logging
    from densephrases import Options
    from densephrases_b200 import runtime as R
    from densephrases_b200 import synthetic_dump as SD
    from oracle.mips_ref import ref_search
    nlist, per_doc = 32, 128
    ntotal = SD.write_synthetic_dump(str(tmp_path), f"start/{nlist}_flat_OPQ96", 128 * 300 + 77, nlist, per_doc, seed=1234)
    assert ntotal == 128 * 300
    o = Options()
    o.add_model_options(); o.add_index_options(); o.add_retrieval_options(); o.add_data_options()
    args = o.parse(["--dump_dir", str(tmp_path), "--index_name", f"start/{nlist}_flat_OPQ96", "--cuda"])
    mips = R.load_phrase_index(args, ignore_logging=True)
    assert mips.index.ntotal == ntotal and mips.is_pq and mips.index.nprobe == 256
    rng = np.random.default_rng(1234)                       # MIPS derives the OPQ matrix from opq_seed the same way
    A = np.linalg.qr(rng.standard_normal((768, 768)))[0].astype(np.float32)
    assert np.array_equal(mips.index.opq_matrix(), A)
    lens = SD.uniform_list_lengths(ntotal, nlist)
    ref = oracle.RefIndex(A, oracle.gen_pq(1234), lens, centroids=oracle.gen_centroids(1234, 0, nlist), seed=1234)
    qr = np.random.default_rng(5)
    pick = qr.integers(0, ntotal - 8, 5)
    vs, _ = ref.reconstruct(pick)
    ve, _ = ref.reconstruct(pick + qr.integers(0, 4, 5))
    query = np.concatenate([vs @ A, ve @ A], 1).astype(np.float64) + 0.05 * qr.standard_normal((5, 1536))
    outs = mips.search(query, q_texts=["q"] * 5, top_k=10, aggregate=True)
    want = ref_search(ref, mips.idx_f, mips.doc_groups, query, top_k=10, nprobe=256, aggregate=True, agg_strat='opt1')
    compare(outs, want)
    assert mips.stage_seconds['batches'] == 1 and mips.stage_seconds['mips'] > 0
    rec = mips.doc_groups['7']
    assert rec['context'][rec['word2char_start'][5]:rec['word2char_end'][5]] in SD._WORDS
