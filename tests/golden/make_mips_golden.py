"""Golden outputs of the UNMODIFIED reference phrase stage: /root/reference/densephrases/index.py `MIPS.search`
(search_dense :189-218, get_idxs :124-141, search_phrase :220-422 PQ + in-RAM metadata branch, aggregate_results :424-448,
adjust :167-176, decompress_meta :106-122) executed in the build container.

The reference module is loaded by file path; its unavailable imports are replaced by inert stand-ins (h5py, faiss, spacy: never
called on this branch) or adapters (blosc.decompress -> densephrases_b200.artifacts.blosc_decompress;
densephrases.utils.eval_utils -> the reference's own file).  `MIPS.__init__` (faiss.read_index ...) is bypassed with __new__ and the
attributes it would set are filled from the synthetic corpus of tests/test_mips.py::build and the CPU oracle index
(index.search / reconstruct).  Everything else -- every line of the methods above -- is the reference's code.

    python tests/golden/make_mips_golden.py   ->   tests/golden/mips_search.json"""
import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference/densephrases"


def load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def main():
    from densephrases_b200 import artifacts
    from oracle import ivfpq_ref as oracle
    from tests import test_mips

    for name in ("h5py", "faiss", "spacy", "spacy.lang"):
        sys.modules[name] = types.ModuleType(name)
    blosc = types.ModuleType("blosc")
    blosc.decompress = artifacts.blosc_decompress
    sys.modules["blosc"] = blosc
    en = types.ModuleType("spacy.lang.en")
    en.English = type("English", (), {})
    sys.modules["spacy.lang.en"] = en
    sys.modules.setdefault("ujson", json)
    for pkg in ("densephrases", "densephrases.utils"):
        sys.modules[pkg] = types.ModuleType(pkg)
    load("densephrases.utils.eval_utils", os.path.join(REF, "utils", "eval_utils.py"))
    ref_index = load("ref_densephrases_index", os.path.join(REF, "index.py"))

    doc_groups, idx_f, _, ref, query = test_mips.build(oracle)
    packed = {}
    for k, g in doc_groups.items():                       # the on-disk form decompress_meta expects (compress_metadata.py:32-53)
        packed[k] = {f: artifacts.blosc_compress(np.asarray(g[f]).tobytes(), typesize=1) for f in ("word2char_start", "word2char_end", "f2o_start")}
        packed[k].update(context=artifacts.blosc_compress(g["context"].encode("utf-8")), title=g["title"],
                         dtypes={f: np.asarray(g[f]).dtype for f in ("word2char_start", "word2char_end", "f2o_start")})

    class FakeFaissIndex:
        ntotal, d = ref.ntotal, ref.d

        def search(self, x, k):
            return ref.search(x, k, 8)

    def reconst_fn(i):
        v, found = ref.reconstruct(np.array([i], dtype=np.int64))
        if not found[0]:
            raise RuntimeError("id not found")          # faiss raises; the reference catches and substitutes zeros (index.py:287-288)
        return v[0]

    def fresh():
        m = ref_index.MIPS.__new__(ref_index.MIPS)
        m.index, m.reconst_fn, m.idx_f, m.doc_groups = FakeFaissIndex(), reconst_fn, idx_f, packed
        m.R = torch.FloatTensor(ref.A.reshape(ref.d, ref.d))
        m.max_idx, m.device, m.cuda, m.num_docs_list, m.offset, m.scale = 1e9, torch.device("cpu"), False, [], None, None
        return m

    out = {"nprobe": 8, "top_k": 5, "configs": []}
    for aggregate, agg in [(False, "opt1"), (True, "opt1"), (True, "opt2"), (True, "opt3"), (True, "opt4")]:
        res = fresh().search(query, q_texts=["q"] * len(query), nprobe=8, top_k=5, aggregate=aggregate, agg_strat=agg, return_idxs=True)
        rows = [[{"context": r["context"], "title": r["title"], "doc_idx": int(r["doc_idx"]), "start_pos": int(r["start_pos"]),
                  "end_pos": int(r["end_pos"]), "start_idx": int(r["start_idx"]), "end_idx": int(r["end_idx"]), "score": float(r["score"]),
                  "answer": r["answer"], "start_vec_sum": float(np.sum(r["start_vec"])), "end_vec_sum": float(np.sum(r["end_vec"]))}
                 for r in rs] for rs in res]
        out["configs"].append({"aggregate": aggregate, "agg_strat": agg, "results": rows})
    # get_idxs on labels outside [0, ntotal): the reference logs, clips and carries on (index.py:128-133)
    I = np.array([[-1, 0, ref.ntotal - 1, ref.ntotal, ref.ntotal + 7], [5, -3, 17, 10 ** 12, 1]], dtype=np.int64)
    doc, word = fresh().get_idxs(I)
    out["get_idxs"] = {"I": I.tolist(), "doc": np.asarray(doc).tolist(), "word": np.asarray(word).tolist()}
    path = os.path.join(ROOT, "tests", "golden", "mips_search.json")
    json.dump(out, open(path, "w"), ensure_ascii=True)
    print(path, os.path.getsize(path), [len(r) for r in out["configs"][0]["results"]])

    # ---- the phrase-dump branch (index.py:246-273): no in-RAM metadata -> h5py groups serve metadata AND int8 token vectors ----
    recs = test_mips.dump_records(doc_groups)

    class Group(dict):                                      # h5py.Group stand-in: datasets by name (numpy arrays), .attrs
        def __init__(self, rec):
            super().__init__({f: rec[f] for f in ("start", "word2char_start", "word2char_end", "f2o_start")})
            self.attrs = {"context": rec["context"], "title": rec["title"]}

    class File(dict):                                       # h5py.File stand-in
        def __init__(self, path, mode):
            super().__init__({k: Group(r) for k, r in recs.items()})

        def close(self):
            pass
    sys.modules["h5py"].File = File

    def fresh_dump():
        m = fresh()
        m.doc_groups, m.phrase_dump_dir = None, "/nonexistent/phrase.hdf5"      # not a directory: one dump file (index.py:98-99)
        return m
    out2 = {"nprobe": 8, "top_k": 5, "configs": []}
    for aggregate, agg, ridx in [(False, "opt1", True), (True, "opt1", False), (True, "opt2", True), (True, "opt4", False)]:
        res = fresh_dump().search(query, q_texts=["q"] * len(query), nprobe=8, top_k=5, aggregate=aggregate, agg_strat=agg, return_idxs=ridx)
        rows = [[dict({"context": r["context"], "title": r["title"], "doc_idx": int(r["doc_idx"]), "start_pos": int(r["start_pos"]),
                       "end_pos": int(r["end_pos"]), "start_idx": int(r["start_idx"]), "end_idx": int(r["end_idx"]), "score": float(r["score"]),
                       "answer": r["answer"]},
                      **({"start_vec_sum": float(np.sum(r["start_vec"])), "end_vec_sum": float(np.sum(r["end_vec"]))} if ridx else {}))
                 for r in rs] for rs in res]
        out2["configs"].append({"aggregate": aggregate, "agg_strat": agg, "return_idxs": ridx, "results": rows})
    path = os.path.join(ROOT, "tests", "golden", "mips_search_hdf5.json")
    json.dump(out2, open(path, "w"), ensure_ascii=True)
    print(path, os.path.getsize(path), [len(r) for r in out2["configs"][0]["results"]])


if __name__ == "__main__":
    main()
