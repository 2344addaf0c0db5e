"""CPU (gloo, world_size 2): host-side logic of the list-range sharded search -- shard_ranges cuts, the one
all-gather exchange (gather_and_merge) and the canonical (score desc, scan position asc) merge -- checked against the
unsharded oracle.  Per-shard partial results are produced by the oracle restricted to the shard's lists."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.helpers import assert_topk_equal, near_queries, opq_matrix


def test_shard_ranges_balance_by_bytes():
    from densephrases_b200.sharded import shard_ranges
    rng = np.random.default_rng(0)
    lens = np.exp(rng.normal(0, 0.5, 4096))
    lens = (lens / lens.sum() * 1e7).astype(np.int64)
    for world in (1, 2, 4, 8):
        rs = shard_ranges(lens, world)
        assert rs[0][0] == 0 and rs[-1][1] == len(lens) and all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
        loads = np.array([lens[a:b].sum() for a, b in rs], dtype=np.float64)
        assert loads.max() / loads.mean() < 1.01
    (a0, a1), (b0, b1) = shard_ranges(np.array([5, 0, 0, 7]), 2)
    assert a0 == 0 and a1 == b0 and b1 == 4 and 1 <= a1 <= 3


def test_query_slice_and_protocol_choice():
    """Slices of the query-split protocol: contiguous, P = ceil(n / W) rows each, zero rows pad the tail; their concatenation cut to
    n rows is the batch.  use_query_split: large batches over a moderate number of lists only."""
    from densephrases_b200.sharded import query_slice, use_query_split
    x = torch.arange(7 * 3, dtype=torch.float32).reshape(7, 3) + 1
    for world in (1, 2, 3, 8):
        parts = [query_slice(x, r, world) for r in range(world)]
        per = (7 + world - 1) // world
        assert all(p.shape == (per, 3) and p.is_contiguous() for p in parts)
        cat = torch.cat(parts)
        assert torch.equal(cat[:7], x) and not cat[7:].any()
    assert use_query_split(1024, 8, 65536) and use_query_split(1024, 2, 65536)          # C4
    assert not use_query_split(128, 8, 1048576) and not use_query_split(64, 8, 4096)     # C5 (few queries, 1M lists); small batches
    assert not use_query_split(1024, 1, 65536)


def numpy_merge(Dg, Ig, Gg, k):
    Dg, Ig, Gg = Dg.numpy(), Ig.numpy(), Gg.numpy().astype(np.int64) & 0xFFFFFFFF
    nsh, n, _ = Dg.shape
    D = np.full((n, k), np.float32(-3.4028234663852886e38), dtype=np.float32)
    I = np.full((n, k), -1, dtype=np.int64)
    for q in range(n):
        ent = [(-float(Dg[s, q, r]), int(Gg[s, q, r]), int(Ig[s, q, r]), Dg[s, q, r]) for s in range(nsh) for r in range(k) if Ig[s, q, r] >= 0]
        ent.sort(key=lambda e: (e[0], e[1]))
        for i, e in enumerate(ent[:k]):
            D[q, i], I[q, i] = e[3], e[2]
    return torch.from_numpy(D), torch.from_numpy(I)


def _fkey(f):
    b = np.asarray(f, dtype=np.float32).view(np.uint32).astype(np.uint64)
    return np.where(b & 0x80000000, (~b) & 0xFFFFFFFF, b | 0x80000000)


def _fkey_inv(k):
    k = np.asarray(k, dtype=np.uint64)
    b = np.where(k & 0x80000000, k & 0x7FFFFFFF, (~k) & 0xFFFFFFFF).astype(np.uint32)
    return b.view(np.float32)


def _worker2(rank, world, port, out):
    """Both exchanges of densephrases_b200.sharded.sharded_search over gloo, local work done by the oracle."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from densephrases_b200.sharded import shard_ranges, sharded_search
    from oracle import ivfpq_ref as R
    seed, nlist, nprobe, k = 9, 40, 10, 10
    lens = np.random.default_rng(seed).integers(0, 300, nlist).astype(np.int64)
    ix = R.RefIndex(opq_matrix(seed), R.gen_pq(seed), lens, centroids=R.gen_centroids(seed, 0, nlist), seed=seed)
    x = near_queries(ix, 7, 3)
    Dfull, Ifull, keyfull = ix.search(x, k, nprobe, return_key=True)
    lo, hi = shard_ranges(lens, world)[rank]
    xr = ix.rotate(x)
    state = {}

    def coarse_local(xt):
        S = R.np_matmul_nt_seq(xr, ix.centroids()[lo:hi])                  # this shard's centroids only (same sequential-FMA scores)
        keys = np.zeros((len(x), nprobe), dtype=np.uint64)
        for q in range(len(x)):
            order = sorted(range(hi - lo), key=lambda j: (-float(S[q, j]), j))[:nprobe]
            for r, j in enumerate(order):
                keys[q, r] = (_fkey(S[q, j]) << np.uint64(32)) | np.uint64(0xFFFFFFFF - (j + lo))
        return torch.from_numpy(keys.view(np.int64))

    def search_preassigned(keys_g, kk):
        kg = keys_g.numpy().view(np.uint64)                                  # [W, n, nprobe]
        key = np.full((len(x), nprobe), -1, dtype=np.int64)
        for q in range(len(x)):
            allk = sorted((int(v) for v in kg[:, q, :].ravel() if v != 0), reverse=True)[:nprobe]
            key[q, :len(allk)] = [0xFFFFFFFF - (v & 0xFFFFFFFF) for v in allk]
        state["key"] = key
        D, I = ix.search_preassigned(xr, np.where((key >= lo) & (key < hi), key, -1), kk)
        G = np.zeros_like(I)
        for q in range(len(x)):
            starts = np.concatenate([[0], np.cumsum([lens[l] if l >= 0 else 0 for l in key[q]])])
            for r in range(kk):
                if I[q, r] >= 0:
                    l, off = ix.locate(np.array([I[q, r]]))
                    G[q, r] = starts[list(key[q]).index(int(l[0]))] + int(off[0])
        return torch.from_numpy(D), torch.from_numpy(I), torch.from_numpy(G.astype(np.int32))

    def pack(D, I, G):
        ck = (_fkey(D.numpy()) << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - (G.numpy().astype(np.int64) & 0xFFFFFFFF).astype(np.uint64))
        P = np.stack([np.where(I.numpy() >= 0, ck, 0).view(np.int64), I.numpy()], axis=-1)
        return torch.from_numpy(np.ascontiguousarray(P))

    def merge_packed(Pg, kk):
        Pn = Pg.numpy()
        D = np.full((len(x), kk), np.float32(-3.4028234663852886e38), dtype=np.float32)
        I = np.full((len(x), kk), -1, dtype=np.int64)
        for q in range(len(x)):
            ent = sorted(((int(np.uint64(Pn[s, q, r, 0])), int(Pn[s, q, r, 1])) for s in range(Pn.shape[0]) for r in range(kk) if Pn[s, q, r, 0] != 0),
                         reverse=True)[:kk]
            for i, (ckey, lab) in enumerate(ent):
                D[q, i], I[q, i] = _fkey_inv(ckey >> 32), lab
        return torch.from_numpy(D), torch.from_numpy(I)

    Dm, Im = sharded_search(torch.from_numpy(x), k, world, None, coarse_local, search_preassigned, pack, merge_packed)

    # ---- the query-split protocol on the same batch: this rank assigns only its slice of the queries, over ALL lists ----
    from densephrases_b200.sharded import sharded_search_qsplit
    R_ = 768 + 2 * nprobe
    qstate = {}

    def coarse_split(xl):
        xl = xl.numpy()
        xrl = ix.rotate(xl)
        S = R.np_matmul_nt_seq(xrl, ix.centroids())
        rec = np.zeros((len(xl), R_), dtype=np.float32)
        for q in range(len(xl)):
            order = sorted(range(nlist), key=lambda j: (-float(S[q, j]), j))[:nprobe]
            rec[q, :768] = xrl[q]
            rec[q, 768:768 + nprobe] = np.array(order, dtype=np.int32).view(np.float32)
            rec[q, 768 + nprobe:] = S[q, order]
        return torch.from_numpy(rec)

    def search_assigned(rec_g, kk):
        rec = rec_g.numpy()
        assert rec.shape == (len(x), R_)                                    # padded rows of the last slice were cut off
        xr_g = np.ascontiguousarray(rec[:, :768])
        key = np.ascontiguousarray(rec[:, 768:768 + nprobe]).view(np.int32).astype(np.int64)
        qstate["key"], qstate["xr"] = key, xr_g
        state["key"] = key
        D, I = ix.search_preassigned(xr_g, np.where((key >= lo) & (key < hi), key, -1), kk)
        G = np.zeros_like(I)
        for q in range(len(x)):
            starts = np.concatenate([[0], np.cumsum([lens[l] if l >= 0 else 0 for l in key[q]])])
            for r in range(kk):
                if I[q, r] >= 0:
                    l, off = ix.locate(np.array([I[q, r]]))
                    G[q, r] = starts[list(key[q]).index(int(l[0]))] + int(off[0])
        return torch.from_numpy(D), torch.from_numpy(I), torch.from_numpy(G.astype(np.int32))

    Dq, Iq = sharded_search_qsplit(torch.from_numpy(x), k, world, rank, None, coarse_split, search_assigned, pack, merge_packed)
    if rank == 0:
        np.savez(out, D=Dm.numpy(), I=Im.numpy(), Dfull=Dfull, Ifull=Ifull, key=state["key"], keyfull=keyfull,
                 Dq=Dq.numpy(), Iq=Iq.numpy(), keyq=qstate["key"], xrq=qstate["xr"], xr=xr)
    dist.destroy_process_group()


def test_gloo_world2_two_exchange_protocol(tmp_path, oracle):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "merged2.npz")
    mp.spawn(_worker2, args=(2, port, out), nprocs=2, join=True)
    g = np.load(out)
    assert np.array_equal(g["key"], g["keyfull"])                       # sharded coarse quantizer == unsharded probe selection
    assert_topk_equal(g["D"], g["I"], g["Dfull"], g["Ifull"], "two-exchange sharded search vs unsharded")
    # query-split protocol (7 queries over 2 ranks: slices of 4 and 3 + one padded row)
    assert np.array_equal(g["keyq"], g["keyfull"]) and np.array_equal(g["xrq"].view(np.int32), g["xr"].view(np.int32))
    assert_topk_equal(g["Dq"], g["Iq"], g["Dfull"], g["Ifull"], "query-split sharded search vs unsharded")


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from densephrases_b200.sharded import gather_and_merge, shard_ranges
    from oracle import ivfpq_ref as R
    seed, nlist, nprobe, k = 5, 32, 12, 10
    lens = np.random.default_rng(seed).integers(0, 300, nlist).astype(np.int64)
    ix = R.RefIndex(opq_matrix(seed), R.gen_pq(seed), lens, centroids=R.gen_centroids(seed, 0, nlist), seed=seed)
    x = near_queries(ix, 9, 3)
    Dfull, Ifull, key = ix.search(x, k, nprobe, return_key=True)
    lo, hi = shard_ranges(lens, world)[rank]
    xr = ix.rotate(x)
    key_local = np.where((key >= lo) & (key < hi), key, -1)
    D, I = ix.search_preassigned(xr, key_local, k)
    # canonical scan position of each hit: prefix of the probed list lengths (ALL probes, global) + offset
    G = np.zeros_like(I)
    for q in range(len(x)):
        starts = np.concatenate([[0], np.cumsum([lens[l] if l >= 0 else 0 for l in key[q]])])
        for r in range(k):
            if I[q, r] >= 0:
                l, off = ix.locate(np.array([I[q, r]]))
                G[q, r] = starts[list(key[q]).index(int(l[0]))] + int(off[0])
    Dm, Im = gather_and_merge(torch.from_numpy(D), torch.from_numpy(I), torch.from_numpy(G.astype(np.int32)), k, world, merge_fn=numpy_merge)
    if rank == 0:
        np.savez(out, D=Dm.numpy(), I=Im.numpy(), Dfull=Dfull, Ifull=Ifull)
    dist.destroy_process_group()


def test_gloo_world2_sharded_merge_equals_unsharded(tmp_path, oracle):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "merged.npz")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    g = np.load(out)
    assert_topk_equal(g["D"], g["I"], g["Dfull"], g["Ifull"], "sharded vs unsharded")


class _OracleShard(object):
    """IvfPqIndex-shaped stand-in for one list-range shard, backed by the CPU oracle: labels outside the shard reconstruct to
    zeros / found 0 and contribute 0 to window scores, exactly like dph_index_reconstruct_batch / dph_index_window_scores."""

    def __init__(self, ref, lo, hi):
        self.ref, self.lo, self.hi = ref, lo, hi
        self.ntotal, self.d, self.nlist = ref.ntotal, ref.d, ref.nlist

    def opq_matrix(self):
        return self.ref.A

    def reconstruct_batch(self, ids):
        v, f = self.ref.reconstruct(ids)
        l, _ = self.ref.locate(ids)
        mine = (l >= self.lo) & (l < self.hi)
        return np.where(mine[:, None], v, 0).astype(np.float32), (f.astype(bool) & mine).astype(np.uint8)

    def window_scores(self, q, first_ids, L):
        out = np.zeros((len(first_ids), L), dtype=np.float32)
        xq = self.ref.rotate(q)                                           # <q, A^T v> == <A q, v>
        for j in range(L):
            v, _ = self.reconstruct_batch(np.asarray(first_ids) + j)
            out[:, j] = (xq * v).sum(1)
        return out


def _worker3(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from densephrases_b200.sharded import ShardedIvfPq, shard_ranges
    from oracle import ivfpq_ref as R
    seed, nlist = 4, 24
    lens = np.random.default_rng(seed).integers(0, 200, nlist).astype(np.int64)
    ref = R.RefIndex(opq_matrix(seed), R.gen_pq(seed), lens, centroids=R.gen_centroids(seed, 0, nlist), seed=seed)
    lo, hi = shard_ranges(lens, world)[rank]
    sh = ShardedIvfPq(nlist, rank=rank, world=world, local=_OracleShard(ref, lo, hi))
    ids = np.concatenate([np.random.default_rng(1).integers(0, ref.ntotal, 40), [-1, ref.ntotal + 5]])
    v, f = sh.reconstruct_batch(ids)
    q = near_queries(ref, 6, 2)
    first = np.random.default_rng(2).integers(0, ref.ntotal - 12, 6)
    w = sh.window_scores(q, first, 10)
    if rank == 0:
        vr, fr = ref.reconstruct(ids)
        xq = ref.rotate(q)
        wr = np.stack([(xq * ref.reconstruct(first + j)[0]).sum(1) for j in range(10)], 1)
        np.savez(out, v=v, f=f, vr=vr, fr=fr, w=w, wr=wr)
    dist.destroy_process_group()


def test_gloo_world2_sharded_reconstruct_and_window_scores(tmp_path, oracle):
    """MIPS over a sharded index: every label lives on one shard, the all-reduce sum of the per-shard answers is the unsharded
    answer (reconstruct rows, found flags, phrase-window scores)."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "recon.npz")
    mp.spawn(_worker3, args=(2, port, out), nprocs=2, join=True)
    g = np.load(out)
    assert np.array_equal(g["v"].view(np.int32), g["vr"].view(np.int32)) and np.array_equal(g["f"], g["fr"])
    assert np.allclose(g["w"], g["wr"], atol=1e-5)
