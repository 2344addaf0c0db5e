"""Index builder (densephrases_b200/build_index.py, the OPQ96 / IVF / PQ96 recipe of build_phrase_index.py:96-150): a TRAINED index
on clustered synthetic vectors, searched by the oracle (CPU) and by the CUDA path (GPU), must retrieve the true maximum-inner-product
neighbours -- the end-to-end property the synthetic random-code indexes cannot show."""
import numpy as np
import pytest


def clustered(n, seed, n_clusters=40, spread=0.35):
    rng = np.random.default_rng(seed)
    centres = rng.standard_normal((n_clusters, 768)).astype(np.float32)
    a = rng.integers(0, n_clusters, n)
    return (centres[a] + spread * rng.standard_normal((n, 768))).astype(np.float32)


@pytest.fixture(scope="module")
def built():
    from densephrases_b200.build_index import build_index
    x = clustered(6000, 1)
    return x, build_index(x[:3000], x, nlist=16, seed=5, niter_opq=2, niter_km=6, niter_pq=4)


def test_trained_index_structure(built):
    x, ix = built
    A = ix["A"]
    assert np.abs(A @ A.T - np.eye(768)).max() < 1e-4                       # OPQ matrix is a rotation
    assert ix["list_len"].sum() == len(x) and (ix["list_len"] > 0).sum() >= 12
    assert sorted(ix["ids"].tolist()) == list(range(len(x)))                # ids = arange + offset (build_phrase_index.py:149-150)
    assert ix["codes"].shape == (len(x), 96) and ix["pq"].shape == (96, 256, 8)


def test_trained_index_recall_with_oracle(built, oracle):
    x, ix = built
    ref = oracle.RefIndex(ix["A"], ix["pq"], ix["list_len"], centroids=ix["centroids"], codes=ix["codes"], ids=ix["ids"])
    rng = np.random.default_rng(3)
    q = x[rng.integers(0, len(x), 40)] + 0.05 * rng.standard_normal((40, 768)).astype(np.float32)
    D, I = ref.search(q, 50, nprobe=8)
    ip = q.astype(np.float64) @ x.astype(np.float64).T
    exact = np.argsort(-ip, axis=1)[:, :10]
    top1 = np.mean([exact[i, 0] == I[i, 0] for i in range(40)])
    # cluster-mates are near-ties (their inner products differ by less than the PQ error), so judge by VALUE and by recall in a wider list
    quality = np.mean([ip[i, I[i, :10]].mean() / ip[i, exact[i]].mean() for i in range(40)])
    recall_at_50 = np.mean([len(set(I[i]) & set(exact[i])) / 10 for i in range(40)])
    assert top1 > 0.9 and quality > 0.97 and recall_at_50 > 0.6, (top1, quality, recall_at_50)
    assert np.abs(D[:, 0] - ip[np.arange(40), I[:, 0]]).max() < 0.15 * np.abs(D[:, 0]).max()      # ADC score ~ true inner product (short training: 3000 points, 4 PQ iterations)
    v, found = ref.reconstruct(ix["ids"][:200])                              # PQ reconstruction error is small against the data scale
    back = v @ ix["A"]
    err = np.linalg.norm(back - x[ix["ids"][:200]], axis=1) / np.linalg.norm(x[ix["ids"][:200]], axis=1)
    assert found.all() and err.mean() < 0.3, err.mean()      # isotropic within-cluster noise is what PQ cannot capture


@pytest.mark.gpu
def test_trained_index_on_gpu_matches_oracle(built, oracle):
    from densephrases_b200 import IvfPqIndex
    from tests.helpers import assert_topk_equal
    x, ix = built
    ref = oracle.RefIndex(ix["A"], ix["pq"], ix["list_len"], centroids=ix["centroids"], codes=ix["codes"], ids=ix["ids"])
    gpu = IvfPqIndex.from_arrays(ix["A"], ix["centroids"], ix["pq"], ix["list_len"], ix["codes"], ix["ids"])
    gpu.nprobe = 8
    q = x[:48] + 0.05 * np.random.default_rng(9).standard_normal((48, 768)).astype(np.float32)
    D, I = gpu.search(q, 10)
    Dr, Ir = ref.search(q, 10, 8)
    assert_topk_equal(D, I, Dr, Ir, "trained index")
    assert (I[:, 0] == np.arange(48)).mean() > 0.9                          # the perturbed vector retrieves itself
