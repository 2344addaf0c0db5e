"""IvfPqIndex: Python handle over the C-ABI index (include/dph_b200.h).

Mirrors the slice of the faiss Python API that /root/reference/densephrases/index.py uses on the hot path:
``search(x, k) -> (D, I)`` (index.py:200), ``reconstruct`` (index.py:31,286,296), ``ntotal``, ``d``, ``nprobe``
(index.py:33,53,62), the OPQ matrix (index.py:32).  numpy in -> numpy out through the C ABI with host buffers;
torch CUDA tensors in -> torch CUDA tensors out (device pointers, asynchronous on the current stream)."""
import ctypes as C

import numpy as np

from . import _lib as L


def _np_ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class IvfPqIndex:
    def __init__(self, nlist, d=768, M=96, nbits=8, device=0):
        self._h = C.c_void_p()
        L.check(L.lib().dph_index_create(C.byref(self._h), d, nlist, M, nbits, device))
        self.device = device

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and L is not None and getattr(L, "_lib", None) is not None:      # interpreter shutdown: the module may already be gone
            L._lib.dph_index_free(h)

    @classmethod
    def from_arrays(cls, A, centroids, pq, list_len, codes, ids=None, device=0, shard=None):
        """Build from the arrays a faiss IndexPreTransform(OPQ)->IndexIVFPQ file holds (index.py:30)."""
        ix = cls(len(list_len), device=device)
        ix.set_opq(A)
        ix.set_centroids(centroids)
        ix.set_pq(pq)
        if shard is not None:
            ix.set_shard(*shard)
        ix.set_lists(list_len, codes, ids)
        return ix

    # ---- construction -----------------------------------------------------------------------
    def set_opq(self, A):
        A = np.ascontiguousarray(A, dtype=np.float32)
        assert A.shape == (self.d, self.d)
        L.check(L.lib().dph_index_set_opq(self._h, _np_ptr(A), L.MEM_HOST))

    def set_centroids(self, Cm):
        Cm = np.ascontiguousarray(Cm, dtype=np.float32)
        assert Cm.shape == (self.nlist, self.d)
        L.check(L.lib().dph_index_set_centroids(self._h, _np_ptr(Cm), L.MEM_HOST))

    def set_pq(self, pq):
        pq = np.ascontiguousarray(pq, dtype=np.float32)
        assert pq.shape == (96, 256, 8)
        L.check(L.lib().dph_index_set_pq(self._h, _np_ptr(pq), L.MEM_HOST))

    def gen_centroids(self, seed, sigma=0.5):
        L.check(L.lib().dph_index_gen_centroids(self._h, seed, sigma))

    def gen_pq(self, seed, sigma=0.25):
        L.check(L.lib().dph_index_gen_pq(self._h, seed, sigma))

    def set_shard(self, lo, hi):
        L.check(L.lib().dph_index_set_shard(self._h, lo, hi))

    def set_lists(self, list_len, codes, ids=None):
        list_len = np.ascontiguousarray(list_len, dtype=np.int64)
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        if ids is not None:
            ids = np.ascontiguousarray(ids, dtype=np.int64)
        L.check(L.lib().dph_index_set_lists(self._h, _np_ptr(list_len), _np_ptr(codes), None if ids is None else _np_ptr(ids)))

    def set_lists_synthetic(self, list_len, seed):
        list_len = np.ascontiguousarray(list_len, dtype=np.int64)
        L.check(L.lib().dph_index_set_lists_synthetic(self._h, _np_ptr(list_len), seed))

    # ---- attributes ---------------------------------------------------------------------------
    @property
    def ntotal(self):
        return L.lib().dph_index_ntotal(self._h)

    @property
    def ntotal_local(self):
        return L.lib().dph_index_ntotal_local(self._h)

    @property
    def d(self):
        return L.lib().dph_index_d(self._h)

    @property
    def nlist(self):
        return L.lib().dph_index_nlist(self._h)

    @property
    def nprobe(self):
        return L.lib().dph_index_nprobe(self._h)

    @nprobe.setter
    def nprobe(self, v):
        L.check(L.lib().dph_index_set_nprobe(self._h, int(v)))

    def set_scan_mode(self, mode):
        L.check(L.lib().dph_index_set_scan_mode(self._h, mode))

    def last_used_pair_mode(self):
        return bool(L.lib().dph_index_last_used_pair_mode(self._h))

    def last_group_size(self):
        """Queries per shared-memory gather in the last search: 1 (fp32 LUT), 2 (pair-packed) or 4 (quad-packed)."""
        return int(L.lib().dph_index_last_group_size(self._h))

    def set_coarse_tc(self, on):
        L.check(L.lib().dph_index_set_coarse_tc(self._h, int(bool(on))))

    def set_stream(self, cuda_stream_ptr):
        L.check(L.lib().dph_index_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def opq_matrix(self):
        A = np.empty((self.d, self.d), dtype=np.float32)
        L.check(L.lib().dph_index_get_opq(self._h, _np_ptr(A), L.MEM_HOST))
        return A

    def set_profile(self, on):
        L.check(L.lib().dph_index_set_profile(self._h, int(bool(on))))

    def last_scan_ms(self):
        ms = C.c_float(0)
        L.check(L.lib().dph_index_last_scan_ms(self._h, C.byref(ms)))
        return ms.value

    def profile_scan_ms(self):
        n = L.lib().dph_index_profile_count(self._h)
        out = np.zeros(max(n, 1), dtype=np.float32)
        L.check(L.lib().dph_index_profile_scan_ms(self._h, _np_ptr(out), n))
        return out[:n]

    @property
    def device_bytes(self):
        return L.lib().dph_index_device_bytes(self._h)

    # ---- search -------------------------------------------------------------------------------
    def search(self, x, k):
        """numpy [n,d] -> (D [n,k] f32, I [n,k] i64) numpy   |   torch cuda [n,d] -> torch cuda (D, I)."""
        if isinstance(x, np.ndarray):
            x = np.ascontiguousarray(x, dtype=np.float32)
            n = x.shape[0]
            D = np.empty((n, k), dtype=np.float32)
            I = np.empty((n, k), dtype=np.int64)
            L.check(L.lib().dph_index_search(self._h, _np_ptr(x), n, k, _np_ptr(D), _np_ptr(I), L.MEM_HOST))
            return D, I
        import torch
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
        n = x.shape[0]
        D = torch.empty((n, k), dtype=torch.float32, device=x.device)
        I = torch.empty((n, k), dtype=torch.int64, device=x.device)
        self.set_stream(torch.cuda.current_stream(x.device).cuda_stream)
        L.check(L.lib().dph_index_search(self._h, x.data_ptr(), n, k, D.data_ptr(), I.data_ptr(), L.MEM_DEVICE))
        return D, I

    def search_partial(self, x, k):
        """torch cuda [n,d] -> per-shard (D, I, G) torch cuda; G = canonical scan position (tie-break)."""
        import torch
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
        n = x.shape[0]
        D = torch.empty((n, k), dtype=torch.float32, device=x.device)
        I = torch.empty((n, k), dtype=torch.int64, device=x.device)
        G = torch.empty((n, k), dtype=torch.int32, device=x.device)
        self.set_stream(torch.cuda.current_stream(x.device).cuda_stream)
        L.check(L.lib().dph_index_search_partial(self._h, x.data_ptr(), n, k, D.data_ptr(), I.data_ptr(), G.data_ptr()))
        return D, I, G

    def coarse_local(self, x):
        """torch cuda [n,d] -> int64 [n,nprobe] keys of this shard's best lists (score key << 32 | ~global list id)."""
        import torch
        n = x.shape[0]
        keys = torch.empty((n, self.nprobe), dtype=torch.int64, device=x.device)
        self.set_stream(torch.cuda.current_stream(x.device).cuda_stream)
        L.check(L.lib().dph_index_coarse_local(self._h, x.data_ptr(), n, keys.data_ptr()))
        return keys

    def search_preassigned(self, keys_gathered, k):
        """all-gathered keys [nshards,n,nprobe] -> this shard's partial (D, I, G) for the batch passed to coarse_local."""
        import torch
        W, n, _ = keys_gathered.shape
        D = torch.empty((n, k), dtype=torch.float32, device=keys_gathered.device)
        I = torch.empty((n, k), dtype=torch.int64, device=keys_gathered.device)
        G = torch.empty((n, k), dtype=torch.int32, device=keys_gathered.device)
        self.set_stream(torch.cuda.current_stream(keys_gathered.device).cuda_stream)
        L.check(L.lib().dph_index_search_preassigned(self._h, keys_gathered.data_ptr(), W, n, k, D.data_ptr(), I.data_ptr(), G.data_ptr()))
        return D, I, G

    def coarse_split(self, x_local):
        """torch cuda [n_local,d] (this rank's slice of the batch) -> records [n_local, 768 + 2 nprobe] f32: rotated query, probed
        lists (int32 bits) and coarse scores over ALL lists (dph_index_coarse_split)."""
        import torch
        n = x_local.shape[0]
        rec = torch.empty((n, L.lib().dph_index_record_floats(self._h)), dtype=torch.float32, device=x_local.device)
        self.set_stream(torch.cuda.current_stream(x_local.device).cuda_stream)
        L.check(L.lib().dph_index_coarse_split(self._h, x_local.data_ptr(), n, rec.data_ptr()))
        return rec

    def search_assigned(self, rec, k):
        """all-gathered records [n, 768 + 2 nprobe] (batch order) -> this shard's partial (D, I, G)."""
        import torch
        assert rec.is_cuda and rec.dtype == torch.float32 and rec.is_contiguous()
        n = rec.shape[0]
        D = torch.empty((n, k), dtype=torch.float32, device=rec.device)
        I = torch.empty((n, k), dtype=torch.int64, device=rec.device)
        G = torch.empty((n, k), dtype=torch.int32, device=rec.device)
        self.set_stream(torch.cuda.current_stream(rec.device).cuda_stream)
        L.check(L.lib().dph_index_search_assigned(self._h, rec.data_ptr(), n, k, D.data_ptr(), I.data_ptr(), G.data_ptr()))
        return D, I, G

    def _last(self, which, shape, dtype):
        out = np.empty(shape, dtype=dtype)
        L.check(L.lib().dph_index_copy_last(self._h, which, _np_ptr(out), out.nbytes))
        return out

    def last_flags(self, n):
        return self._last(0, (n,), np.int32)

    def last_probes(self, n):
        return self._last(1, (n, self.nprobe), np.int32)

    def last_coarse(self, n):
        return self._last(2, (n, self.nprobe), np.float32)

    def last_xr(self, n):
        return self._last(3, (n, self.d), np.float32)

    # ---- reconstruct ----------------------------------------------------------------------------
    def reconstruct_batch(self, ids):
        """labels [m] -> (vec [m,d] f32 in ROTATED space, found [m] u8); missing label -> zeros (index.py:287-288)."""
        if isinstance(ids, np.ndarray) or isinstance(ids, (list, tuple)):
            ids = np.ascontiguousarray(ids, dtype=np.int64)
            out = np.empty((len(ids), self.d), dtype=np.float32)
            found = np.empty(len(ids), dtype=np.uint8)
            L.check(L.lib().dph_index_reconstruct_batch(self._h, _np_ptr(ids), len(ids), _np_ptr(out), _np_ptr(found), L.MEM_HOST))
            return out, found
        import torch
        assert ids.is_cuda and ids.dtype == torch.int64
        out = torch.empty((ids.numel(), self.d), dtype=torch.float32, device=ids.device)
        found = torch.empty(ids.numel(), dtype=torch.uint8, device=ids.device)
        self.set_stream(torch.cuda.current_stream(ids.device).cuda_stream)
        L.check(L.lib().dph_index_reconstruct_batch(self._h, ids.data_ptr(), ids.numel(), out.data_ptr(), found.data_ptr(), L.MEM_DEVICE))
        return out, found


def _window_scores(self, q, first_ids, L):
    """q [m,768] f32, first_ids [m] int64 (numpy) -> scores [m,L] f32: <q[i], un-rotated reconstruct(first_ids[i] + l)>;
    labels that are not in the index score 0 (the reference's zero vector, index.py:287-288)."""
    q = np.ascontiguousarray(q, dtype=np.float32)
    first_ids = np.ascontiguousarray(first_ids, dtype=np.int64)
    out = np.empty((len(first_ids), L), dtype=np.float32)
    L_.check(L_.lib().dph_index_window_scores(self._h, _np_ptr(q), _np_ptr(first_ids), len(first_ids), L, _np_ptr(out), L_.MEM_HOST))
    return out


L_ = L
IvfPqIndex.window_scores = _window_scores


def merge_shards(Dg, Ig, Gg, k):
    """all-gathered [nshards,n,k] torch cuda tensors -> (D, I) [n,k]; order score desc, scan position asc."""
    import torch
    nsh, n, kk = Dg.shape
    assert kk == k and Dg.is_contiguous() and Ig.is_contiguous() and Gg.is_contiguous()
    D = torch.empty((n, k), dtype=torch.float32, device=Dg.device)
    I = torch.empty((n, k), dtype=torch.int64, device=Dg.device)
    st = torch.cuda.current_stream(Dg.device).cuda_stream
    L.check(L.lib().dph_merge_shards(Dg.data_ptr(), Ig.data_ptr(), Gg.data_ptr(), nsh, n, k, D.data_ptr(), I.data_ptr(), C.c_void_p(st)))
    return D, I


def pack_topk(D, I, G):
    """(D f32, I i64, G i32) [n,k] cuda -> P int64 [n,k,2] for a single all-gather."""
    import torch
    n, k = D.shape
    P = torch.empty((n, k, 2), dtype=torch.int64, device=D.device)
    st = torch.cuda.current_stream(D.device).cuda_stream
    L.check(L.lib().dph_pack_topk(D.data_ptr(), I.data_ptr(), G.data_ptr(), n, k, P.data_ptr(), C.c_void_p(st)))
    return P


def merge_shards_packed(Pg, k):
    """all-gathered P [nshards,n,k,2] -> (D, I) [n,k]."""
    import torch
    nsh, n = Pg.shape[0], Pg.shape[1]
    D = torch.empty((n, k), dtype=torch.float32, device=Pg.device)
    I = torch.empty((n, k), dtype=torch.int64, device=Pg.device)
    st = torch.cuda.current_stream(Pg.device).cuda_stream
    L.check(L.lib().dph_merge_shards_packed(Pg.data_ptr(), nsh, n, k, D.data_ptr(), I.data_ptr(), C.c_void_p(st)))
    return D, I
