"""Readers for the reference's on-disk artefacts (densephrases_b200/artifacts.py): index.faiss, idx2id.hdf5, meta_compressed.pkl.
faiss / h5py / blosc do not exist in the build container, so these tests pin the readers against the writers of the same module
(both restate the published formats) and against hand-assembled byte strings for the fixed parts of each header."""
import os
import pickle
import struct
import zlib

import numpy as np
import pytest

from densephrases_b200 import artifacts as A


def small_index(rng, nlist=7, lens=(3, 0, 5, 1, 0, 0, 2)):
    ll = np.array(lens, dtype=np.int64)
    return dict(A=rng.standard_normal((768, 768)).astype(np.float32), centroids=rng.standard_normal((nlist, 768)).astype(np.float32),
                pq=rng.standard_normal((96, 256, 8)).astype(np.float32), list_len=ll,
                codes=rng.integers(0, 256, (int(ll.sum()), 96), dtype=np.uint8), ids=rng.permutation(10 ** 6)[:int(ll.sum())].astype(np.int64) + 3 * 10 ** 9)


@pytest.mark.parametrize("variant", ["ilar_full", "ilar_sparse", "ilod"])
def test_faiss_index_round_trip(tmp_path, variant):
    rng = np.random.default_rng(1)
    ix = small_index(rng) if variant != "ilar_sparse" else small_index(rng, 9, (0, 0, 4, 0, 0, 0, 0, 2, 0))
    p = str(tmp_path / "index.faiss")
    A.write_faiss_index(p, nprobe=256, ondisk_payload="merged_index.ivfdata" if variant == "ilod" else None, **ix)
    raw = open(p, "rb").read()
    assert raw[:4] == b"IxPT" and struct.unpack_from("<i", raw, 4)[0] == 768            # fourcc, then d of the index header
    assert (b"sprs" in raw) == (variant == "ilar_sparse") and (b"ilod" in raw) == (variant == "ilod")
    got = A.read_faiss_index(p)
    for k, v in ix.items():
        assert np.array_equal(got[k], v), k
    assert got["nprobe"] == 256 and got["ntotal"] == int(ix["list_len"].sum()) and got["by_residual"] and got["metric"] == 0
    if variant == "ilod":                                                              # payload is looked up NEXT TO the index file (index.py:30)
        os.rename(tmp_path / "merged_index.ivfdata", tmp_path / "moved.ivfdata")
        with pytest.raises(FileNotFoundError):
            A.read_faiss_index(p)


def test_faiss_reader_fails_loudly(tmp_path):
    p = tmp_path / "bad.faiss"
    p.write_bytes(b"IxSQ" + b"\0" * 64)
    with pytest.raises(ValueError, match="IxSQ"):
        A.read_faiss_index(str(p))
    rng = np.random.default_rng(2)
    good = tmp_path / "index.faiss"
    A.write_faiss_index(str(good), **small_index(rng))
    (tmp_path / "cut.faiss").write_bytes(good.read_bytes()[:-100])
    with pytest.raises(ValueError, match="truncated"):
        A.read_faiss_index(str(tmp_path / "cut.faiss"))


@pytest.mark.parametrize("typesize,shuffle,split,n", [(1, True, False, 5000), (8, True, True, 100000), (8, True, False, 100000), (4, False, True, 33333),
                                                      (8, True, True, 50), (1, False, False, 0), (2, True, True, 70001)])
def test_blosc_frames(typesize, shuffle, split, n):
    rng = np.random.default_rng(n + typesize)
    data = rng.integers(0, 4, n, dtype=np.uint8).tobytes()
    frame = A.blosc_compress(data, typesize=typesize, shuffle=shuffle, split=split)
    ver, verlz, flags, ts, nbytes, blocksize, cbytes = struct.unpack_from("<BBBBIII", frame, 0)
    assert (ver, ts, nbytes, cbytes) == (2, typesize, n, len(frame))
    assert A.blosc_decompress(frame) == data
    if n > 1000:
        assert len(frame) < n                     # really compressed
        with pytest.raises(ValueError):
            A.blosc_decompress(frame[:-7] + b"\0" * 7)
    with pytest.raises(ValueError, match="not supported"):
        A.blosc_decompress(bytes([2, 1, (1 << 5) | 1, 1]) + struct.pack("<III", 1000, 1000, 300) + b"\0" * 284)     # lz4 frame


def test_meta_fields_in_all_three_encodings(tmp_path):
    w = np.arange(0, 900, 3, dtype=np.int32)
    ctx = "Łódź is a city. " * 40
    rec_ref = {"word2char_start": A.blosc_compress(w.tobytes(), typesize=1), "context": A.blosc_compress(ctx.encode("utf-8"), typesize=8, split=True),
               "dtypes": {"word2char_start": w.dtype}, "title": ["T"]}
    p = tmp_path / "meta_compressed.pkl"
    pickle.dump({"0": rec_ref}, open(p, "wb"))
    rec = A.read_meta(str(p))["0"]
    assert np.array_equal(A.decode_meta_field(rec["word2char_start"], rec["dtypes"]["word2char_start"]), w)
    assert A.decode_meta_field(rec["context"]).decode("utf-8") == ctx
    assert np.array_equal(A.decode_meta_field(zlib.compress(w.tobytes()), w.dtype), w)           # this repo's converter
    assert np.array_equal(A.decode_meta_field(w, w.dtype), w)                                    # raw


def same(a, b):
    if isinstance(a, dict):
        return set(a) == set(b) and all(same(a[k], b[k]) for k in a)
    return a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b)


def test_hdf5_round_trip_and_superblock(tmp_path):
    rng = np.random.default_rng(3)
    tree = {str(i * 10 ** 9): {"doc": rng.integers(0, 5000, 1000 + i).astype(np.int32), "word": rng.integers(0, 300, 1000 + i).astype(np.int32)}
            for i in range(40)}                                                                  # 40 links -> 5 symbol nodes under the root B-tree
    p = str(tmp_path / "idx2id.hdf5")
    A.write_hdf5(p, tree)
    raw = open(p, "rb").read()
    assert raw[:8] == b"\x89HDF\r\n\x1a\n" and raw[8] == 0 and raw[13] == 8 and raw[14] == 8
    assert struct.unpack_from("<Q", raw, 40)[0] == len(raw)                                     # end-of-file address
    assert raw.count(b"SNOD") >= 5 + 40 and raw.count(b"HEAP") == 41
    assert same(tree, A.read_idx2id(p))
    tree["misc"] = {"f": rng.standard_normal((3, 5)).astype(np.float32), "be": np.arange(10, dtype=">i8"), "empty": np.zeros(0, np.int32),
                    "deep": {"x": np.arange(4, dtype=np.uint8)}, "tab": rng.integers(-9, 9, (37, 11)).astype(np.int16)}
    A.write_hdf5(p, tree, chunks={"tab": (16, 4)})                                               # chunked + shuffle + deflate
    got = A.read_hdf5(p)
    assert same(tree, got)
    body = bytearray(open(p, "rb").read())                                                       # user block: superblock at 512, base address 512
    struct.pack_into("<Q", body, 24, 512)
    (tmp_path / "user.hdf5").write_bytes(b"\0" * 512 + bytes(body))
    assert same(tree, A.read_hdf5(str(tmp_path / "user.hdf5")))
    with pytest.raises(ValueError, match="not an HDF5"):
        (tmp_path / "junk").write_bytes(b"junk" * 300)
        A.read_hdf5(str(tmp_path / "junk"))


def test_mips_loads_the_reference_file_layout(tmp_path, monkeypatch):
    """MIPS(phrase_dump_dir, index_path, idx2id_path) over index.faiss (merged, on-disk lists) + idx2id.hdf5 + meta_compressed.pkl,
    with the device index replaced by a recorder (CPU test): the arrays handed to ShardedIvfPq.from_arrays (world size 1 = one IvfPqIndex) are the file's."""
    from densephrases_b200 import mips, sharded
    from densephrases_b200.synthetic import make_corpus, make_phrase_index_arrays
    rng = np.random.default_rng(4)
    doc_groups, idx_f, ntotal = make_corpus(6, 1)
    list_len, codes, ids = make_phrase_index_arrays(ntotal, 8, 1)
    ix = dict(A=np.linalg.qr(rng.standard_normal((768, 768)))[0].astype(np.float32), centroids=rng.standard_normal((8, 768)).astype(np.float32),
              pq=rng.standard_normal((96, 256, 8)).astype(np.float32), list_len=list_len, codes=codes, ids=ids)
    dump = tmp_path / "dump"
    (dump / "phrase").mkdir(parents=True)
    (dump / "start" / "8_flat_OPQ96").mkdir(parents=True)
    index_path = str(dump / "start" / "8_flat_OPQ96" / "index.faiss")
    A.write_faiss_index(index_path, nprobe=1, ondisk_payload="merged_index.ivfdata", **ix)
    A.write_hdf5(str(dump / "start" / "8_flat_OPQ96" / "idx2id.hdf5"), {k: {t: np.asarray(v[t]) for t in ("doc", "word")} for k, v in idx_f.items()})
    packed = {}
    for k, g in doc_groups.items():
        packed[k] = {f: A.blosc_compress(np.asarray(g[f]).tobytes(), typesize=1) for f in ("word2char_start", "word2char_end", "f2o_start")}
        packed[k].update(context=A.blosc_compress(g["context"].encode("utf-8")), title=g["title"],
                         dtypes={f: np.asarray(g[f]).dtype for f in ("word2char_start", "word2char_end", "f2o_start")})
    pickle.dump(packed, open(dump / "meta_compressed.pkl", "wb"))

    class Recorder:
        d, nprobe = 768, 1

        def __init__(self, *arrays):
            self.arrays = arrays
            self.ntotal = int(np.sum(arrays[3]))

        def opq_matrix(self):
            return self.arrays[0]

        reconstruct_batch = None

    monkeypatch.setattr(sharded.ShardedIvfPq, "from_arrays", staticmethod(lambda *a, **kw: Recorder(*a)))
    m = mips.MIPS(str(dump / "phrase"), index_path, str(dump / "start" / "8_flat_OPQ96" / "idx2id.hdf5"), cuda=False)
    for got, want in zip(m.index.arrays, (ix["A"], ix["centroids"], ix["pq"], list_len, codes, ids)):
        assert np.array_equal(got, want)
    assert m.index.nprobe == 256 and m.is_pq and same({k: {t: np.asarray(v[t]) for t in ("doc", "word")} for k, v in idx_f.items()}, m.idx_f)
    k0 = sorted(doc_groups)[0]
    meta = m.decompress_meta(k0)
    assert meta["context"] == doc_groups[k0]["context"] and np.array_equal(meta["f2o_start"], doc_groups[k0]["f2o_start"])
    assert np.allclose(m.R.numpy(), ix["A"])


# ---- cross-checks against the real libraries: run automatically wherever faiss / h5py / blosc can be imported (none of them exists in
# ---- the build container or on the GPU box, so they are skipped there; DESIGN.md keeps the readers marked "unverified" until one runs)
def test_faiss_reader_and_search_oracle_against_real_faiss(tmp_path):
    """A tiny IndexPreTransform(OPQMatrix, IndexIVFPQ(IndexFlatIP)) built and written by faiss itself (build_phrase_index.py:113-116,
    142): (1) artifacts.read_faiss_index parses the file faiss wrote, (2) the CPU oracle (oracle/ivfpq_ref.c) reproduces faiss'
    own search results on it -- the check that would lift 'parity unpinned' (labels equal, scores within 1e-3)."""
    faiss = pytest.importorskip("faiss")
    if "IndexIVFPQ" not in vars(faiss):
        pytest.skip("only this repo's importable `faiss` stub is present (eval_phrase_retrieval.py:12 imports the name), not the library")
    from oracle import ivfpq_ref as R
    rng = np.random.default_rng(0)
    d, nlist, n = 768, 8, 4000
    xb = rng.standard_normal((n, d)).astype(np.float32)
    quantizer = faiss.IndexFlatIP(d)
    sub = faiss.IndexIVFPQ(quantizer, d, nlist, 96, 8, faiss.METRIC_INNER_PRODUCT)
    opq = faiss.OPQMatrix(d, 96)
    opq.niter = 2
    index = faiss.IndexPreTransform(opq, sub)
    index.train(xb)
    index.add_with_ids(xb, np.arange(n, dtype=np.int64) * 7 + 5)
    p = str(tmp_path / "index.faiss")
    faiss.write_index(index, p)
    got = A.read_faiss_index(p)
    assert got["ntotal"] == n and got["by_residual"] and got["metric"] == 0 and got["A"].shape == (d, d) and got["pq"].shape == (96, 256, 8)
    assert np.array_equal(got["A"].ravel(), faiss.vector_to_array(opq.A))
    ref = R.RefIndex(got["A"], got["pq"], got["list_len"], centroids=got["centroids"], codes=got["codes"], ids=got["ids"])
    xq = rng.standard_normal((9, d)).astype(np.float32)
    faiss.extract_index_ivf(index).nprobe = 4
    Df, If = index.search(xq, 10)
    Dr, Ir = ref.search(xq, 10, 4)
    assert np.array_equal(If, Ir) and np.abs(Df - Dr).max() < 1e-3


def test_hdf5_groups_datasets_and_string_attributes_round_trip(tmp_path):
    """The phrase-dump subset (one group per document, int8 / int datasets, `context` / `title` string attributes:
    embed_utils.py:233-246) written by this module's writer -- variable-length UTF-8 strings in a global heap collection like h5py's
    `g.attrs[k] = "text"`, and fixed-length strings -- and read back lazily."""
    rng = np.random.default_rng(5)
    tree = {str(d): {"start": rng.integers(-128, 128, (T, 16), dtype=np.int8), "f2o_start": np.arange(T, dtype=np.int64)} for d, T in ((3, 5), (40, 1), (41, 0))}
    attrs = {"3": {"context": "Paris is the capital. [PAR] caf\u00e9 \u2713 " * 200, "title": "Paris"}, "40": {"context": "", "title": b"fixed"},
             "41": {"context": "x", "title": "T"}}
    p = str(tmp_path / "0-1.hdf5")
    A.write_hdf5(p, tree, attrs=attrs)
    h = A.open_hdf5(p)
    root = h.children(h.root_addr)
    assert sorted(root) == sorted(tree)
    for k, addr in root.items():
        got, members = h.attributes(addr), h.children(addr)
        assert got["context"] == attrs[k]["context"]
        assert got["title"] == (attrs[k]["title"].decode() if isinstance(attrs[k]["title"], bytes) else attrs[k]["title"])
        assert np.array_equal(h.dataset(members["start"]), tree[k]["start"]) and h.dataset(members["start"]).dtype == np.int8
    assert h.attributes(h.root_addr) == {}
    assert np.array_equal(A.read_hdf5(p)["3"]["f2o_start"], tree["3"]["f2o_start"])      # attributes do not disturb the eager reader


def test_phrase_dump_reader_against_h5py(tmp_path):
    """Auto-enabled where h5py exists: a phrase-dump-like file written by h5py itself, read by the native reader (PhraseDump)."""
    h5py = pytest.importorskip("h5py")
    from densephrases_b200.phrase_dump import _NativeFile
    rng = np.random.default_rng(1)
    p = str(tmp_path / "0-1.hdf5")
    want = {}
    with h5py.File(p, "w") as f:
        for d in (0, 17):
            g = f.create_group(str(d))
            want[str(d)] = {"start": rng.integers(-128, 128, (9, 768), dtype=np.int8), "f2o_start": np.arange(9), "word2char_start": np.arange(9) * 3,
                            "word2char_end": np.arange(9) * 3 + 2, "context": "some context caf\u00e9 " * 30, "title": f"title {d}"}
            for name in ("start", "f2o_start", "word2char_start", "word2char_end"):
                g.create_dataset(name, data=want[str(d)][name])
            g.attrs["context"], g.attrs["title"] = want[str(d)]["context"], want[str(d)]["title"]
    nf = _NativeFile(p)
    for k, w in want.items():
        assert nf.has(k)
        rec = nf.group(k)
        assert rec["context"] == w["context"] and rec["title"] == w["title"]
        for name in ("start", "f2o_start", "word2char_start", "word2char_end"):
            assert np.array_equal(np.asarray(rec[name]), w[name])


def test_hdf5_reader_against_h5py(tmp_path):
    h5py = pytest.importorskip("h5py")
    rng = np.random.default_rng(0)
    p = str(tmp_path / "idx2id.hdf5")
    want = {}
    with h5py.File(p, "w") as f:                                  # build_phrase_index.py:268-276
        for off in (0, 1_000_000_000):
            g = f.create_group(str(off))
            want[str(off)] = {"doc": rng.integers(0, 1000, 777).astype(np.int32), "word": rng.integers(0, 300, 777).astype(np.int32)}
            g.create_dataset("doc", data=want[str(off)]["doc"])
            g.create_dataset("word", data=want[str(off)]["word"])
    got = A.read_idx2id(p)
    assert set(got) == set(want)
    for k in want:
        assert np.array_equal(got[k]["doc"], want[k]["doc"]) and np.array_equal(got[k]["word"], want[k]["word"])


def test_blosc_decoder_against_blosc():
    blosc = pytest.importorskip("blosc")
    rng = np.random.default_rng(0)
    for arr in (rng.integers(0, 5000, 3000).astype(np.int32), np.arange(100000, dtype=np.int64), np.zeros(0, np.int32)):
        frame = blosc.compress(arr.tobytes(), typesize=arr.dtype.itemsize, cname="zlib")       # compress_metadata.py:32-53
        assert A.blosc_decompress(frame) == arr.tobytes()
