// scan.cu -- the PQ asymmetric-distance scan over the probed inverted lists and the top-k merge.
//
// Replaces faiss IVFPQScanner::scan_codes + the result heap behind self.index.search(...) at
// /root/reference/densephrases/index.py:200.
//
// Work decomposition: the plan kernel linearises all in-shard (query, probe, 32-vector block) triples in
// canonical scan order; scan CTA c of G takes the contiguous range [T c/G, T (c+1)/G).  One persistent CTA
// per SM (the per-query LUT fills shared memory); inside a CTA each warp owns one 32-vector block per round,
// lane l <-> vector l.
//
// FAST (one query per gather, fp32 LUT).  Shared-memory gathers, not HBM, co-limit this scan (one 4-byte LUT gather per code
// byte).  The code blocks are stored lane-rotated (common.cuh: dph_blk_addr) and the LUT is stored as three [256][64] tables
// (32 sub-quantizers + 31 wrap copies per row), so that at step t lane l reads word (l + t%32) of row `code byte` of table t/32:
// the 32 lanes always hit 32 different banks -> every LDS is one conflict-free wavefront regardless of the code values.  One
// PRMT builds the address (code<<8 | lane*4), so a lookup is PRMT + LDS + FADD.  The rotated summation order differs from faiss'
// m-ascending order by at most eps (plan kernel), so the scores are used as a FILTER: each CTA keeps its best k+slack by filter
// score, the merge kernel re-scores the survivors in canonical order (bit-exact with the oracle) and PROVES that nothing that was
// dropped could have been in the top-k (T_k - max drop threshold > 2 eps); otherwise the query is flagged and re-run through
// EXACT mode.
//
// PAIR (scan_pair_kernel, further down): two queries that probe the same list share every gather through int16-packed quantised
// LUTs -- the default whenever lists are shared by the batch (the reference's nprobe = 256).
//
// EXACT: canonical m-ascending fp32 sum for every code (bank conflicts and all) -- fallback/cross-check.
#include "index_internal.cuh"
#include "select.cuh"

struct ScanArgs {
    const uint8_t* codes; const long long* qpre; const DphSeg* segs; const int* nseg; const DphWork* work;
    const float* lut_canon;
    unsigned* gthr; unsigned long long* cand; const long long* cand_off; int* cand_cnt;
    long long n; int nprobe; int keep;
};

#define NT DPH_SCAN_THREADS
#define NW DPH_SCAN_WARPS
#define SMEM_LUT_FAST (DPH_LUT_SCAN_FLOATS * 4)                 // 196608
#define SMEM_LUT_EXACT (DPH_LUT_CANON_FLOATS * 4 + NW * DPH_BLK_BYTES)   // 98304 + 49152

struct ScanShared {   // tail of the dynamic shared memory
    unsigned long long cbuf[DPH_CAND_CAP];
    DphSeg segtab[DPH_SEG_SMEM];
    SelectScratch sc;
    int cnt; unsigned thr; int base; int ndone; int ndone_snap; int pad[3];
};

__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
// TMA-engine bulk prefetch of one contiguous code block into L2 (no register / shared-memory cost).
__device__ __forceinline__ void l2_prefetch_block(const void* p) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(p), "n"(DPH_BLK_BYTES) : "memory");
}

extern __shared__ __align__(1024) unsigned char dph_smem[];
// The dynamic shared memory window of a kernel without static shared memory starts at this shared-space address on
// sm_100 (1 KB is reserved by the system); probed once at library init (dph_scan_setup_attrs), a mismatch is fatal.
#define DPH_DYN_SMEM_BASE 0x400

// One FAST lookup: PRMT builds (cta window bits | code << 8 | lane*4); the table base, the step offset and the window
// base are folded into the LDS immediate -> PRMT + LDS + FADD per code byte.
template <int IMM> __device__ __forceinline__ float lds_imm(unsigned addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(IMM));
    return v;
}
// The four running sums live in two float2 registers and advance with packed FADD2 (add.rn.f32x2: two independent IEEE fp32 adds per
// instruction) -- the same per-component summation order as four scalar accumulators, half the FP32-pipe instructions.
template <int T0> __device__ __forceinline__ void fast_word(unsigned wv, unsigned y, float2& a01, float2& a23) {
    constexpr int TB = DPH_DYN_SMEM_BASE + (T0 >> 5) * 65536 + (T0 & 31) * 4;   // T0 % 4 == 0: the 4 bytes share a table
    const float v0 = lds_imm<TB + 0>(__byte_perm(wv, y, 0x7504));
    const float v1 = lds_imm<TB + 4>(__byte_perm(wv, y, 0x7514));
    const float v2 = lds_imm<TB + 8>(__byte_perm(wv, y, 0x7524));
    const float v3 = lds_imm<TB + 12>(__byte_perm(wv, y, 0x7534));
    a01 = __fadd2_rn(a01, make_float2(v0, v1));
    a23 = __fadd2_rn(a23, make_float2(v2, v3));
}
template <int C> __device__ __forceinline__ void fast_chunk(const uint4& v, unsigned y, float2& a01, float2& a23) {
    fast_word<C * 16 + 0>(v.x, y, a01, a23);
    fast_word<C * 16 + 4>(v.y, y, a01, a23);
    fast_word<C * 16 + 8>(v.z, y, a01, a23);
    fast_word<C * 16 + 12>(v.w, y, a01, a23);
}

// Segment cursor: blocks are visited in increasing order, so a forward-only cursor finds the segment of work block b
// (relative to the query's first block).  The segment fields are cached in registers and re-read only when a
// segment boundary is crossed (every ~48 rounds at C2 sizes), all lanes reading the same table entry (broadcast).
struct SegCursor {
    int seg; unsigned wrel, wend; const uint4* base; int len; unsigned gstart; float dis0;
    __device__ __forceinline__ void init() { seg = -1; wrel = 0; wend = 0; base = nullptr; len = 0; gstart = 0; dis0 = 0.f; }
    __device__ __forceinline__ void seek(const DphSeg* __restrict__ tab, unsigned b, const uint8_t* codes) {
        if (b < wend) return;
        do { seg++; } while (b >= tab[seg].wend);
        const uint4 u0 = *reinterpret_cast<const uint4*>(tab + seg);
        const uint4 u1 = *(reinterpret_cast<const uint4*>(tab + seg) + 1);
        base = reinterpret_cast<const uint4*>(codes + (long long)(((unsigned long long)u0.y << 32) | u0.x) * DPH_BLK_BYTES);
        len = (int)u0.z; gstart = u0.w; dis0 = __uint_as_float(u1.x); wrel = u1.y; wend = u1.z;
    }
    __device__ __forceinline__ const uint4* ptr(unsigned b) const { return base + (size_t)(b - wrel) * (DPH_BLK_BYTES / 16); }
};

__device__ __forceinline__ void block_compact(ScanShared* sh, int keep, unsigned* gthr_q) {
    const int tid = threadIdx.x;
    const int n = sh->cnt;
    unsigned long long* cb = sh->cbuf;
    unsigned long long pivot = block_radix_select([&](int i) { return cb[i]; }, n, keep, &sh->sc);
    unsigned long long mine[DPH_CAND_CAP / NT];
#pragma unroll
    for (int e = 0; e < DPH_CAND_CAP / NT; e++) { int i = tid + e * NT; mine[e] = (i < n) ? cb[i] : 0ull; }
    __syncthreads();
    if (tid == 0) sh->cnt = 0;
    __syncthreads();
#pragma unroll
    for (int e = 0; e < DPH_CAND_CAP / NT; e++)
        if (mine[e] >= pivot && mine[e] != 0ull) { int p = atomicAdd(&sh->cnt, 1); cb[p] = mine[e]; }
    if (tid == 0) {
        unsigned t = (unsigned)(pivot >> 32);
        unsigned old = atomicMax(gthr_q, t);
        sh->thr = t > old ? t : old;
        sh->ndone_snap = sh->ndone;
    }
    __syncthreads();
}

template <int MODE>
__global__ void __launch_bounds__(NT, 1) scan_kernel(ScanArgs a) {
    unsigned char* const smem = dph_smem;
    constexpr int LUT_BYTES = (MODE == DPH_SCAN_FAST) ? SMEM_LUT_FAST : SMEM_LUT_EXACT;
    ScanShared* sh = reinterpret_cast<ScanShared*>(smem + LUT_BYTES);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long T = a.work->total_blocks;
    const long long G = gridDim.x, c = blockIdx.x;
    const long long g0 = T * c / G, g1 = T * (c + 1) / G;
    if (g0 >= g1) return;
    long long lo = 0, hi = a.n;
    while (hi - lo > 1) { long long mid = (lo + hi) >> 1; if (a.qpre[mid] <= g0) lo = mid; else hi = mid; }
    long long q = lo, g = g0;
    const unsigned ywin = (((unsigned)__cvta_generic_to_shared(dph_smem)) & 0xFF000000u) | ((unsigned)lane * 4u);

    while (g < g1) {
        while (a.qpre[q + 1] <= g) q++;
        const long long qstart = a.qpre[q];
        const long long gend = (a.qpre[q + 1] < g1) ? a.qpre[q + 1] : g1;
        const unsigned b0 = (unsigned)(g - qstart), b1 = (unsigned)(gend - qstart);
        __syncthreads();
        {
            const int nsg0 = a.nseg[q];
            if (nsg0 <= DPH_SEG_SMEM) {
                const uint4* src = reinterpret_cast<const uint4*>(a.segs + q * a.nprobe);
                uint4* dst = reinterpret_cast<uint4*>(sh->segtab);
                for (int i = tid; i < nsg0 * 2; i += NT) dst[i] = __ldg(src + i);
            }
        }
        if (tid == 0) { sh->cnt = 0; sh->ndone = 0; sh->ndone_snap = 0; sh->thr = *((volatile unsigned*)(a.gthr + q)); }
        __syncthreads();

        // segment table of this query: shared memory when it fits (always for nprobe <= 256), else global
        const DphSeg* tab = (a.nseg[q] <= DPH_SEG_SMEM) ? sh->segtab : (a.segs + q * a.nprobe);
        SegCursor cc, pc;                                 // consume / L2-prefetch cursors
        cc.init(); pc.init();
        unsigned b = b0 + warp;                           // next block this warp consumes
        unsigned bp = b0 + warp;                          // next block this warp prefetches into L2
        uint4 nxt[6];
        int n_len = 0; unsigned n_gstart = 0, n_j0 = 0; float n_dis0 = 0.f;
        // the first code blocks are requested BEFORE the LUT is staged: their HBM latency overlaps the staging
#pragma unroll 1
        for (int r = 0; r < DPH_L2_PREFETCH_ROUNDS && bp < b1; r++, bp += NW) {
            pc.seek(tab, bp, a.codes);
            if (lane == 0) l2_prefetch_block(pc.ptr(bp));
        }
        bool more = b < b1;
        if (more) {
            cc.seek(tab, b, a.codes);
            const uint4* p = cc.ptr(b) + lane;
#pragma unroll
            for (int c6 = 0; c6 < 6; c6++) nxt[c6] = ldg_stream(p + c6 * 32);
            n_len = cc.len; n_gstart = cc.gstart; n_dis0 = cc.dis0; n_j0 = (b - cc.wrel) * 32u;
        }
        // ---- stage this query's LUT into shared memory (twelve 16-byte loads in flight per thread) ----
        {
            const float4* src = reinterpret_cast<const float4*>(a.lut_canon + (size_t)q * DPH_LUT_CANON_FLOATS);
            constexpr int PER_T = DPH_LUT_CANON_FLOATS / 4 / NT;      // 12
            constexpr int LB = 12;
            static_assert(PER_T % LB == 0, "LUT staging batches");
            if (MODE == DPH_SCAN_FAST) {
                // canonical rows [seg][code][32] -> scan rows [seg][code][64]: words 0..31 = the row, words 32..62 = its first 31
                // entries again (the wrap copies that let lane l read word l + t without a modulo), word 63 unused.  One thread
                // moves one float4; a quarter-warp writes 128 contiguous bytes of one row -> conflict-free.
                float* dstf = reinterpret_cast<float*>(smem);
#pragma unroll 1
                for (int i0 = tid; i0 < DPH_LUT_CANON_FLOATS / 4; i0 += NT * LB) {
                    float4 v[LB];
#pragma unroll
                    for (int e = 0; e < LB; e++) v[e] = __ldg(src + i0 + e * NT);
#pragma unroll
                    for (int e = 0; e < LB; e++) {
                        const int i = i0 + e * NT;
                        const int row = i >> 3, w = (i & 7) * 4;              // row = seg * 256 + code
                        float* r = dstf + row * 64;
                        *reinterpret_cast<float4*>(r + w) = v[e];
                        if (w < 28) *reinterpret_cast<float4*>(r + 32 + w) = v[e];
                        else { r[60] = v[e].x; r[61] = v[e].y; r[62] = v[e].z; r[63] = 0.0f; }
                    }
                }
            } else {
                float4* dst = reinterpret_cast<float4*>(smem);
#pragma unroll 8
                for (int i = tid; i < DPH_LUT_CANON_FLOATS / 4; i += NT) dst[i] = __ldg(src + i);
            }
        }
        __syncthreads();
        bool counted = false;
        while (true) {      // epochs: run until the candidate buffer may overflow or this warp is out of blocks, then meet
            while (more) {
                if (*((volatile int*)&sh->cnt) > DPH_CAND_CAP - NT) break;     // every warp adds <= 32 per round after this check
                uint4 cur[6];
#pragma unroll
                for (int c6 = 0; c6 < 6; c6++) cur[c6] = nxt[c6];
                const int c_len = n_len; const unsigned c_gstart = n_gstart, c_j0 = n_j0; const float c_dis0 = n_dis0;
                if (bp < b1) {
                    pc.seek(tab, bp, a.codes);
                    if (lane == 0) l2_prefetch_block(pc.ptr(bp));
                    bp += NW;
                }
                b += NW;
                more = b < b1;
                if (more) {
                    cc.seek(tab, b, a.codes);
                    const uint4* p = cc.ptr(b) + lane;
#pragma unroll
                    for (int c6 = 0; c6 < 6; c6++) nxt[c6] = ldg_stream(p + c6 * 32);
                    n_len = cc.len; n_gstart = cc.gstart; n_dis0 = cc.dis0; n_j0 = (b - cc.wrel) * 32u;
                }
                float score;
                if (MODE == DPH_SCAN_FAST) {
                    float2 a01 = make_float2(c_dis0, 0.f), a23 = make_float2(0.f, 0.f);
                    fast_chunk<0>(cur[0], ywin, a01, a23);
                    fast_chunk<1>(cur[1], ywin, a01, a23);
                    fast_chunk<2>(cur[2], ywin, a01, a23);
                    fast_chunk<3>(cur[3], ywin, a01, a23);
                    fast_chunk<4>(cur[4], ywin, a01, a23);
                    fast_chunk<5>(cur[5], ywin, a01, a23);
                    score = (a01.x + a01.y) + (a23.x + a23.y);
                } else {
                    unsigned char* stage = smem + DPH_LUT_CANON_FLOATS * 4 + warp * DPH_BLK_BYTES;
                    const float* lutc = reinterpret_cast<const float*>(smem);
#pragma unroll
                    for (int c6 = 0; c6 < 6; c6++) *reinterpret_cast<uint4*>(stage + c6 * 512 + lane * 16) = cur[c6];
                    __syncwarp();
                    float dis = c_dis0;
#pragma unroll 8
                    for (int m = 0; m < DPH_M; m++) dis += lutc[DPH_LUTC_IDX(m, (int)stage[dph_blk_addr(lane, m)])];
                    score = dis;
                    __syncwarp();
                }
                const unsigned j = c_j0 + lane;
                const unsigned sk = dph_fkey(score);
                const bool pass = ((int)j < c_len) && (sk >= *((volatile unsigned*)&sh->thr));
                const unsigned pm = __ballot_sync(0xffffffffu, pass);
                if (pm) {
                    int basep = 0;
                    if (lane == 0) basep = atomicAdd(&sh->cnt, __popc(pm));
                    basep = __shfl_sync(0xffffffffu, basep, 0);
                    if (pass) {
                        int p = basep + __popc(pm & ((1u << lane) - 1u));
                        if (p < DPH_CAND_CAP) sh->cbuf[p] = ((unsigned long long)sk << 32) | (unsigned long long)(0xFFFFFFFFu - (c_gstart + j));
                    }
                }
            }
            if (!more && !counted) { counted = true; if (lane == 0) atomicAdd(&sh->ndone, 1); }
            __syncthreads();
            if (sh->cnt > a.keep) block_compact(sh, a.keep, a.gthr + q);
            else {
                if (tid == 0) {
                    unsigned gt = *((volatile unsigned*)(a.gthr + q));
                    if (gt > sh->thr) sh->thr = gt;
                    sh->ndone_snap = sh->ndone;
                }
                __syncthreads();
            }
            if (sh->ndone_snap == NW) break;
        }
        // ---- publish this CTA's candidates for query q ----
        __syncthreads();
        const int cnt = sh->cnt;
        if (tid == 0) sh->base = atomicAdd(a.cand_cnt + q, cnt);
        __syncthreads();
        {
            const long long off = a.cand_off[q], cap = a.cand_off[q + 1] - off;
            const int basep = sh->base;
            for (int i = tid; i < cnt; i += NT)
                if (basep + i < cap) a.cand[off + basep + i] = sh->cbuf[i];
        }
        g = gend;
        q++;
    }
}

// =================================================================================================
// PAIR mode: when a list is probed by several queries of the batch (C2: 64 x 256 probes over 4096 lists = 4 per list),
// two of them share every gather: their LUTs are quantised to 10-bit integers (plan: lutq_kernel) and packed into one
// 32-bit word (query a in the low half, query b in the high half), so ONE PRMT + LDS + IADD advances both running sums
// (96 x 682 < 2^16: the halves cannot carry into each other).  Per (query, byte) that halves the code bytes read from HBM,
// the shared-memory gathers and the instructions.  The integer sums are exact; the only error is the quantisation
// (<= 0.5 step per entry), which the plan folds into eps, so the same proof + exact canonical re-scoring applies.
// Work: the plan inverts probes into per-list query groups and pairs them; an item = (list, query pair), linearised by
// blocks; a CTA takes a contiguous block range, rebuilding the packed LUT (192 KB, from L2) at every item boundary.
// =================================================================================================
struct PairScanArgs {
    const uint8_t* codes; const long long* blk_off; const int* list_len;
    const int* pl_cnt; const int* pl_off; const unsigned long long* units; const unsigned* entries; const DphPairWork* work; int* next_unit;
    const unsigned short* lutq; const float2* qparams; const float* cd; const unsigned* gdense;
    unsigned* gthr; unsigned long long* cand; const long long* cand_off; int* cand_cnt;
    long long list_lo, list_hi; int nprobe; int keep;
    const DphUnit* udesc;          // quad mode: resolved unit records (plan)
    unsigned one;                  // the constant 1, as a run-time value (imad_add)
};
#define PCAP 1536
struct PairShared {
    unsigned long long cbuf[2][PCAP];
    SelectScratch sc;
    int cnt[2]; unsigned thr[2]; int base[2]; int ndone; int ndone_snap; int unit;
};
template <int IMM> __device__ __forceinline__ unsigned lds_imm_u32(unsigned addr) {
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(IMM));
    return v;
}
template <int T0> __device__ __forceinline__ void pair_word(unsigned wv, unsigned y, unsigned& a0, unsigned& a1, unsigned& a2, unsigned& a3) {
    constexpr int TB = DPH_DYN_SMEM_BASE + (T0 >> 5) * 65536 + (T0 & 31) * 4;
    a0 += lds_imm_u32<TB + 0>(__byte_perm(wv, y, 0x7504));
    a1 += lds_imm_u32<TB + 4>(__byte_perm(wv, y, 0x7514));
    a2 += lds_imm_u32<TB + 8>(__byte_perm(wv, y, 0x7524));
    a3 += lds_imm_u32<TB + 12>(__byte_perm(wv, y, 0x7534));
}
template <int C> __device__ __forceinline__ void pair_chunk(const uint4& v, unsigned y, unsigned& a0, unsigned& a1, unsigned& a2, unsigned& a3) {
    pair_word<C * 16 + 0>(v.x, y, a0, a1, a2, a3);
    pair_word<C * 16 + 4>(v.y, y, a0, a1, a2, a3);
    pair_word<C * 16 + 8>(v.z, y, a0, a1, a2, a3);
    pair_word<C * 16 + 12>(v.w, y, a0, a1, a2, a3);
}
template <int CAP>
__device__ __forceinline__ void compact_buffer(unsigned long long* cb, int* cnt, unsigned* thr, int keep, unsigned* gthr_q, SelectScratch* sc) {
    const int tid = threadIdx.x;
    const int n = *cnt;
    unsigned long long pivot = block_radix_select([&](int i) { return cb[i]; }, n, keep, sc);
    constexpr int PER = (CAP + NT - 1) / NT;
    unsigned long long mine[PER];
#pragma unroll
    for (int e = 0; e < PER; e++) { int i = tid + e * NT; mine[e] = (i < n) ? cb[i] : 0ull; }
    __syncthreads();
    if (tid == 0) *cnt = 0;
    __syncthreads();
#pragma unroll
    for (int e = 0; e < PER; e++)
        if (mine[e] >= pivot && mine[e] != 0ull) { int p = atomicAdd(cnt, 1); cb[p] = mine[e]; }
    if (tid == 0) {
        unsigned t = (unsigned)(pivot >> 32);
        unsigned old = atomicMax(gthr_q, t);
        *thr = t > old ? t : old;
    }
    __syncthreads();
}
__device__ __forceinline__ void warp_append(bool pass, unsigned long long key, unsigned long long* cb, int* cnt, int cap, int lane) {
    const unsigned pm = __ballot_sync(0xffffffffu, pass);
    if (pm) {
        int basep = 0;
        if (lane == 0) basep = atomicAdd(cnt, __popc(pm));
        basep = __shfl_sync(0xffffffffu, basep, 0);
        if (pass) { int p = basep + __popc(pm & ((1u << lane) - 1u)); if (p < cap) cb[p] = key; }
    }
}

__global__ void __launch_bounds__(NT, 1) scan_pair_kernel(PairScanArgs a) {
    unsigned char* const smem = dph_smem;
    PairShared* sh = reinterpret_cast<PairShared*>(smem + SMEM_LUT_FAST);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int total_units = a.work->total_units;
    const unsigned segb = (unsigned)a.work->per;
    const unsigned ywin = (((unsigned)__cvta_generic_to_shared(dph_smem)) & 0xFF000000u) | ((unsigned)lane * 4u);

    while (true) {
        // pull the next unit of the queue: consecutive units are the items of one list segment, so they run at the same time on
        // different CTAs and share the segment's code blocks through L2
        if (tid == 0) sh->unit = atomicAdd(a.next_unit, 1);
        __syncthreads();
        const int u = sh->unit;
        if (u >= total_units) break;
        const unsigned long long ud = a.units[u];
        const long long l = (long long)(ud & 0xFFFFFFFFull);
        const int it = (int)((ud >> 32) & 0xFFFFull);
        const int len = a.list_len[l];
        const unsigned nb = (unsigned)((len + 31) >> 5);
        const unsigned bi0 = (unsigned)(ud >> 48) * segb;
        const unsigned bend = (nb - bi0 < segb) ? nb : bi0 + segb;
        const int e0 = a.pl_off[l] + 2 * it;
        const bool has_b = (2 * it + 1) < a.pl_cnt[l];
        const unsigned ea = a.entries[e0], eb = has_b ? a.entries[e0 + 1] : ea;
        const long long qa = ea >> 10, qb = eb >> 10;
        const int ra = (int)(ea & 1023u), rb = (int)(eb & 1023u);
        __syncthreads();
        {   // ---- packed LUT: word = qa entry | qb entry << 16, same [3][256][64] scan layout ----
            const uint4* A4 = reinterpret_cast<const uint4*>(a.lutq + (size_t)qa * DPH_LUT_SCAN_FLOATS);
            const uint4* B4 = reinterpret_cast<const uint4*>(a.lutq + (size_t)qb * DPH_LUT_SCAN_FLOATS);
            uint4* dst = reinterpret_cast<uint4*>(smem);
            for (int i = tid; i < DPH_LUT_SCAN_FLOATS / 8; i += NT) {
                const uint4 va = __ldg(A4 + i);
                uint4 vb = make_uint4(0u, 0u, 0u, 0u);
                if (has_b) vb = __ldg(B4 + i);
                uint4 o0, o1;
                o0.x = (va.x & 0xFFFFu) | (vb.x << 16); o0.y = (va.x >> 16) | (vb.x & 0xFFFF0000u);
                o0.z = (va.y & 0xFFFFu) | (vb.y << 16); o0.w = (va.y >> 16) | (vb.y & 0xFFFF0000u);
                o1.x = (va.z & 0xFFFFu) | (vb.z << 16); o1.y = (va.z >> 16) | (vb.z & 0xFFFF0000u);
                o1.z = (va.w & 0xFFFFu) | (vb.w << 16); o1.w = (va.w >> 16) | (vb.w & 0xFFFF0000u);
                dst[2 * i] = o0; dst[2 * i + 1] = o1;
            }
        }
        if (tid == 0) {
            sh->cnt[0] = 0; sh->cnt[1] = 0; sh->ndone = 0; sh->ndone_snap = 0;
            sh->thr[0] = *((volatile unsigned*)(a.gthr + qa));
            sh->thr[1] = has_b ? *((volatile unsigned*)(a.gthr + qb)) : 0xFFFFFFFFu;
        }
        __syncthreads();
        const float2 pa = a.qparams[qa], pb = a.qparams[qb];
        const float base_a = a.cd[qa * a.nprobe + ra] + pa.y, base_b = a.cd[qb * a.nprobe + rb] + pb.y;
        const unsigned gs_a = a.gdense[qa * a.nprobe + ra], gs_b = a.gdense[qb * a.nprobe + rb];
        const uint4* lbase = reinterpret_cast<const uint4*>(a.codes + a.blk_off[l] * DPH_BLK_BYTES);

        unsigned b = bi0 + warp, bp = bi0 + warp;
        uint4 nxt[6];
#pragma unroll 1
        for (int r = 0; r < DPH_L2_PREFETCH_ROUNDS && bp < bend; r++, bp += NW)
            if (lane == 0) l2_prefetch_block(lbase + (size_t)bp * (DPH_BLK_BYTES / 16));
        bool more = b < bend;
        if (more) {
            const uint4* p = lbase + (size_t)b * (DPH_BLK_BYTES / 16) + lane;
#pragma unroll
            for (int c6 = 0; c6 < 6; c6++) nxt[c6] = ldg_stream(p + c6 * 32);
        }
        bool counted = false;
        while (true) {
            while (more) {
                if (*((volatile int*)&sh->cnt[0]) > PCAP - NT || *((volatile int*)&sh->cnt[1]) > PCAP - NT) break;
                uint4 cur[6];
#pragma unroll
                for (int c6 = 0; c6 < 6; c6++) cur[c6] = nxt[c6];
                const unsigned bcur = b;
                if (bp < bend) { if (lane == 0) l2_prefetch_block(lbase + (size_t)bp * (DPH_BLK_BYTES / 16)); bp += NW; }
                b += NW;
                more = b < bend;
                if (more) {
                    const uint4* p = lbase + (size_t)b * (DPH_BLK_BYTES / 16) + lane;
#pragma unroll
                    for (int c6 = 0; c6 < 6; c6++) nxt[c6] = ldg_stream(p + c6 * 32);
                }
                unsigned s0 = 0, s1 = 0, s2 = 0, s3 = 0;
                pair_chunk<0>(cur[0], ywin, s0, s1, s2, s3);
                pair_chunk<1>(cur[1], ywin, s0, s1, s2, s3);
                pair_chunk<2>(cur[2], ywin, s0, s1, s2, s3);
                pair_chunk<3>(cur[3], ywin, s0, s1, s2, s3);
                pair_chunk<4>(cur[4], ywin, s0, s1, s2, s3);
                pair_chunk<5>(cur[5], ywin, s0, s1, s2, s3);
                const unsigned tot = (s0 + s1) + (s2 + s3);
                const float sc_a = fmaf(pa.x, (float)(tot & 0xFFFFu), base_a);
                const float sc_b = fmaf(pb.x, (float)(tot >> 16), base_b);
                const unsigned j = bcur * 32u + lane;
                const bool valid = (int)j < len;
                const unsigned ka = dph_fkey(sc_a), kb = dph_fkey(sc_b);
                warp_append(valid && ka >= *((volatile unsigned*)&sh->thr[0]), ((unsigned long long)ka << 32) | (unsigned long long)(0xFFFFFFFFu - (gs_a + j)),
                            sh->cbuf[0], &sh->cnt[0], PCAP, lane);
                warp_append(valid && kb >= *((volatile unsigned*)&sh->thr[1]), ((unsigned long long)kb << 32) | (unsigned long long)(0xFFFFFFFFu - (gs_b + j)),
                            sh->cbuf[1], &sh->cnt[1], PCAP, lane);
            }
            if (!more && !counted) { counted = true; if (lane == 0) atomicAdd(&sh->ndone, 1); }
            __syncthreads();
            if (sh->cnt[0] > a.keep) compact_buffer<PCAP>(sh->cbuf[0], &sh->cnt[0], &sh->thr[0], a.keep, a.gthr + qa, &sh->sc);
            if (sh->cnt[1] > a.keep) compact_buffer<PCAP>(sh->cbuf[1], &sh->cnt[1], &sh->thr[1], a.keep, a.gthr + qb, &sh->sc);
            if (tid == 0) {
                unsigned ga = *((volatile unsigned*)(a.gthr + qa));
                if (ga > sh->thr[0]) sh->thr[0] = ga;
                if (has_b) { unsigned gb = *((volatile unsigned*)(a.gthr + qb)); if (gb > sh->thr[1]) sh->thr[1] = gb; }
                sh->ndone_snap = sh->ndone;
            }
            __syncthreads();
            if (sh->ndone_snap == NW) break;
        }
        // ---- publish both candidate sets ----
        for (int s2i = 0; s2i < (has_b ? 2 : 1); s2i++) {
            const long long q = s2i ? qb : qa;
            const int cnt = sh->cnt[s2i];
            if (tid == 0) sh->base[s2i] = atomicAdd(a.cand_cnt + q, cnt);
            __syncthreads();
            const long long off = a.cand_off[q], cap = a.cand_off[q + 1] - off;
            const int basep = sh->base[s2i];
            for (int i = tid; i < cnt; i += NT)
                if (basep + i < cap) a.cand[off + basep + i] = sh->cbuf[s2i][i];
        }
    }
}


// =================================================================================================
// QUAD mode: FOUR queries that probe the same list share every gather.  Their LUTs are quantised to 8-bit integers
// (plan: lutq_kernel<unsigned char, 255>) and packed into one 32-bit word (query i in byte i), so one PRMT + LDS serves
// four (query, code byte) lookups -- the shared-memory gather pipe, which bounds the pair kernel, does half the work per
// lookup.  The byte lanes would overflow after two additions, so the running sums are kept as TWO 32-bit registers:
//     sraw  = sum of the gathered words, plain 32-bit wrap-around arithmetic
//           = S0 + 2^8 S1 + 2^16 S2 + 2^24 S3   (mod 2^32),        S_i = sum of query i's entries  (<= 96 * 255 < 2^15)
//     accb  = sum of PRMT(word -> [b1, 0, b3, 0]) = S1 + 2^16 S3     (exact: both lanes stay below 2^16)
// and decoded once per vector:  sraw - (accb << 8) = S0 + 2^16 S2 (exact, < 2^32).
// The integer sums are exact; the only error is the quantisation (<= 0.5 step per entry), which the plan folds into eps, so the
// same proof + exact canonical re-scoring applies (merge_kernel).  Work items are (list segment, group of <= 4 probing queries).
// =================================================================================================
#define QCAP 768
struct QuadShared {
    unsigned long long cbuf[4][QCAP];
    SelectScratch sc;
    DphUnit desc[2];                 // current item / next item (fetched with cp.async while the current one is scanned)
    int cnt[4]; unsigned thr[4]; int base[4]; int ndone; int ndone_snap; int unit; int full;
    long long qs[4];                 // the group's queries (shared copy: indexed by thread id in the publish step)
};
// 32-bit add on the FMA pipe: d = a * one + c with `one` a kernel parameter holding 1 (opaque to ptxas, so the multiply-add is not
// folded back into an IADD3).  The quad scan is bound by the ALU pipe (PRMT + IADD3, one warp instruction per two cycles per
// sub-partition, B300_MICROARCH "fma vs alu split"); IMAD issues on the otherwise idle FMA pipe at the same rate.
__device__ __forceinline__ unsigned imad_add(unsigned x, unsigned one, unsigned c) {
    unsigned d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(one), "r"(c));
    return d;
}
// IMADL: how many of the running-sum additions go to the FMA pipe: 0 = none (all IADD3), 1 = the raw sums, 2 = the raw sums and
// half of the odd-byte sums, 3 = all of them.
template <int T0, int IMADL> __device__ __forceinline__ void quad_word(unsigned wv, unsigned y, unsigned one, unsigned (&sr)[4], unsigned (&ab)[4]) {
    constexpr int TB = DPH_DYN_SMEM_BASE + (T0 >> 5) * 65536 + (T0 & 31) * 4;
    constexpr int A = (T0 >> 2) & 1;                  // two accumulator sets, alternating per code word
    const unsigned w0 = lds_imm_u32<TB + 0>(__byte_perm(wv, y, 0x7504));
    const unsigned w1 = lds_imm_u32<TB + 4>(__byte_perm(wv, y, 0x7514));
    const unsigned w2 = lds_imm_u32<TB + 8>(__byte_perm(wv, y, 0x7524));
    const unsigned w3 = lds_imm_u32<TB + 12>(__byte_perm(wv, y, 0x7534));
    // 8-bit entries: the odd bytes are widened per gather (a lane-wise add of two words could carry).  A 7-bit variant that widens
    // once per TWO gathers is 11 % faster (ncu r2d) but doubles eps, and then the exactness proof fails for ~10 % of the queries at
    // C2 (measured); the exact fallback costs far more than it saves.
    if (IMADL >= 1) {
        sr[2 * A] = imad_add(w1, one, imad_add(w0, one, sr[2 * A]));
        sr[2 * A + 1] = imad_add(w3, one, imad_add(w2, one, sr[2 * A + 1]));
    } else {
        sr[2 * A] += w0 + w1;
        sr[2 * A + 1] += w2 + w3;
    }
    const unsigned e0 = __byte_perm(w0, 0u, 0x4341), e1 = __byte_perm(w1, 0u, 0x4341);
    const unsigned e2 = __byte_perm(w2, 0u, 0x4341), e3 = __byte_perm(w3, 0u, 0x4341);
    if (IMADL >= 3 || (IMADL == 2 && A == 1)) {
        ab[2 * A] = imad_add(e1, one, imad_add(e0, one, ab[2 * A]));
        ab[2 * A + 1] = imad_add(e3, one, imad_add(e2, one, ab[2 * A + 1]));
    } else {
        ab[2 * A] += e0 + e1;
        ab[2 * A + 1] += e2 + e3;
    }
}
template <int C, int IMADL> __device__ __forceinline__ void quad_chunk(const uint4& v, unsigned y, unsigned one, unsigned (&sr)[4], unsigned (&ab)[4]) {
    quad_word<C * 16 + 0, IMADL>(v.x, y, one, sr, ab);
    quad_word<C * 16 + 4, IMADL>(v.y, y, one, sr, ab);
    quad_word<C * 16 + 8, IMADL>(v.z, y, one, sr, ab);
    quad_word<C * 16 + 12, IMADL>(v.w, y, one, sr, ab);
}
// append with an overflow latch: the round loop polls ONE flag instead of every buffer's counter
__device__ __forceinline__ void warp_append_latch(bool pass, unsigned long long key, unsigned long long* cb, int* cnt, int* full, int lane) {
    const unsigned pm = __ballot_sync(0xffffffffu, pass);
    if (pm) {
        int basep = 0;
        const int np = __popc(pm);
        if (lane == 0) { basep = atomicAdd(cnt, np); if (basep + np > QCAP - NT) *((volatile int*)full) = 1; }
        basep = __shfl_sync(0xffffffffu, basep, 0);
        if (pass) { int p = basep + __popc(pm & ((1u << lane) - 1u)); if (p < QCAP) cb[p] = key; }
    }
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}

// Item start-up is kept off the critical path (at C2 / C4 list lengths an item is only 30-50 rounds long): the plan resolves every
// unit into one DphUnit record; the record of the NEXT unit is fetched into shared memory with cp.async while the current one is
// scanned (its queue index was requested at the start of the current item); the first code blocks of an item are requested BEFORE the
// packed LUT is built, so the HBM latency overlaps the build; the build reads only the 32 real bytes of every 64-byte table row and
// writes the wrap copy itself (half the L2 traffic).
template <int IMADL>
__global__ void __launch_bounds__(NT, 1) scan_quad_kernel(PairScanArgs a) {
    unsigned char* const smem = dph_smem;
    QuadShared* sh = reinterpret_cast<QuadShared*>(smem + SMEM_LUT_FAST);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int total_units = a.work->total_units;
    const unsigned ywin = (((unsigned)__cvta_generic_to_shared(dph_smem)) & 0xFF000000u) | ((unsigned)lane * 4u);
    const unsigned char* lut8 = reinterpret_cast<const unsigned char*>(a.lutq);
    const unsigned one = a.one;

    if (tid == 0) sh->unit = atomicAdd(a.next_unit, 1);
    __syncthreads();
    {
        const int u = sh->unit;
        if (u >= total_units) return;
        if (tid < (int)(sizeof(DphUnit) / 16)) reinterpret_cast<uint4*>(&sh->desc[0])[tid] = __ldg(reinterpret_cast<const uint4*>(a.udesc + u) + tid);
    }
    __syncthreads();
    int cur = 0;
    while (true) {
        int next_u = 0;
        if (tid == 0) next_u = atomicAdd(a.next_unit, 1);          // consumed after the LUT build (fetch of the next record)
        const DphUnit* dsc = &sh->desc[cur];
        const int len = dsc->len, nq = dsc->nq;
        const unsigned bi0 = dsc->bi0, bend = dsc->bend;
        const uint4* lbase = reinterpret_cast<const uint4*>(a.codes + dsc->blk * DPH_BLK_BYTES);
        unsigned qv[4], gsv[4]; float stepv[4], basev[4];
#pragma unroll
        for (int i = 0; i < 4; i++) { qv[i] = dsc->q[i]; gsv[i] = dsc->gs[i]; basev[i] = dsc->base[i]; stepv[i] = dsc->step[i]; }

        // ---- first code blocks: L2 prefetch + this warp's first block into registers, in flight during the LUT build ----
        unsigned b = bi0 + warp, bp = bi0 + warp;
        uint4 nxt[6];
#pragma unroll 1
        for (int r = 0; r < DPH_L2_PREFETCH_ROUNDS && bp < bend; r++, bp += NW)
            if (lane == 0) l2_prefetch_block(lbase + (size_t)bp * (DPH_BLK_BYTES / 16));
        bool more = b < bend;
        if (more) {
            const uint4* p = lbase + (size_t)b * (DPH_BLK_BYTES / 16) + lane;
#pragma unroll
            for (int c6 = 0; c6 < 6; c6++) nxt[c6] = ldg_stream(p + c6 * 32);
        }
        unsigned g4[4] = {0u, 0u, 0u, 0u};
        if (tid == 0) {
#pragma unroll
            for (int i = 0; i < 4; i++) g4[i] = i < nq ? *((volatile unsigned*)(a.gthr + qv[i])) : 0xFFFFFFFFu;
        }
        {   // ---- packed LUT: byte i of every word = query i's 8-bit entry, [3][256][64] scan layout (words 32..62 = words 0..30) ----
            const unsigned* T0p = reinterpret_cast<const unsigned*>(lut8 + (size_t)qv[0] * DPH_LUT_SCAN_FLOATS);
            const unsigned* T1p = reinterpret_cast<const unsigned*>(lut8 + (size_t)qv[1] * DPH_LUT_SCAN_FLOATS);
            const unsigned* T2p = reinterpret_cast<const unsigned*>(lut8 + (size_t)qv[2] * DPH_LUT_SCAN_FLOATS);
            const unsigned* T3p = reinterpret_cast<const unsigned*>(lut8 + (size_t)qv[3] * DPH_LUT_SCAN_FLOATS);
            uint4* dst = reinterpret_cast<uint4*>(smem);
            // (row, group of four real words): a quarter-warp moves one row.  The empty slots of a short group repeat query 0 (their
            // byte lanes are summed like the others and never pass: threshold +inf), so all four loads are unconditional and a
            // thread keeps 24 of them in flight.
            constexpr int LB = 6;
#pragma unroll 1
            for (int i0 = tid; i0 < 768 * 8; i0 += NT * LB) {
                unsigned va[LB], vb[LB], vc[LB], vd[LB];
#pragma unroll
                for (int e = 0; e < LB; e++) {
                    const int i = i0 + e * NT, src = (i >> 3) * 16 + (i & 7);
                    va[e] = __ldg(T0p + src); vb[e] = __ldg(T1p + src); vc[e] = __ldg(T2p + src); vd[e] = __ldg(T3p + src);
                }
#pragma unroll
                for (int e = 0; e < LB; e++) {
                    const int i = i0 + e * NT, row = i >> 3, cg = i & 7;
                    const unsigned t0 = __byte_perm(va[e], vb[e], 0x5140), t1 = __byte_perm(va[e], vb[e], 0x7362);     // [a0 b0 a1 b1], [a2 b2 a3 b3]
                    const unsigned u0 = __byte_perm(vc[e], vd[e], 0x5140), u1 = __byte_perm(vc[e], vd[e], 0x7362);
                    uint4 o;
                    o.x = __byte_perm(t0, u0, 0x5410); o.y = __byte_perm(t0, u0, 0x7632);
                    o.z = __byte_perm(t1, u1, 0x5410); o.w = __byte_perm(t1, u1, 0x7632);
                    dst[row * 16 + cg] = o;
                    dst[row * 16 + 8 + cg] = o;                        // wrap copy (word 63 = word 31: never read)
                }
            }
        }
        if (tid == 0) {
            sh->ndone = 0; sh->ndone_snap = 0; sh->full = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) { sh->cnt[i] = 0; sh->qs[i] = (long long)qv[i]; sh->thr[i] = g4[i]; }
        }
        __syncthreads();
        if (warp == 0) {          // the next unit's record -> the other descriptor slot (waited for at the end of this item)
            const int nu = __shfl_sync(0xffffffffu, next_u, 0);
            if (lane == 0) sh->unit = nu;
            if (nu < total_units && lane < (int)(sizeof(DphUnit) / 16))
                cp_async16(reinterpret_cast<unsigned char*>(&sh->desc[cur ^ 1]) + lane * 16, reinterpret_cast<const unsigned char*>(a.udesc + nu) + lane * 16);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        unsigned thr0 = sh->thr[0], thr1 = sh->thr[1], thr2 = sh->thr[2], thr3 = sh->thr[3];     // refreshed after every barrier
        // float images of the thresholds (0xFFFFFFFF = "never passes" for the empty slots of a short group -> +inf)
        auto thr_f = [](unsigned t) { return t == 0xFFFFFFFFu ? __int_as_float(0x7f800000) : (t == 0u ? -__int_as_float(0x7f800000) : dph_fkey_inv(t)); };
        float tf0 = thr_f(thr0), tf1 = thr_f(thr1), tf2 = thr_f(thr2), tf3 = thr_f(thr3);

        bool counted = false;
        while (true) {
            while (more) {
                if (*((volatile int*)&sh->full)) break;
                uint4 cur6[6];
#pragma unroll
                for (int c6 = 0; c6 < 6; c6++) cur6[c6] = nxt[c6];
                const unsigned bcur = b;
                if (bp < bend) { if (lane == 0) l2_prefetch_block(lbase + (size_t)bp * (DPH_BLK_BYTES / 16)); bp += NW; }
                b += NW;
                more = b < bend;
                if (more) {
                    const uint4* p = lbase + (size_t)b * (DPH_BLK_BYTES / 16) + lane;
#pragma unroll
                    for (int c6 = 0; c6 < 6; c6++) nxt[c6] = ldg_stream(p + c6 * 32);
                }
                unsigned sr[4] = {0u, 0u, 0u, 0u}, ab[4] = {0u, 0u, 0u, 0u};
                quad_chunk<0, IMADL>(cur6[0], ywin, one, sr, ab);
                quad_chunk<1, IMADL>(cur6[1], ywin, one, sr, ab);
                quad_chunk<2, IMADL>(cur6[2], ywin, one, sr, ab);
                quad_chunk<3, IMADL>(cur6[3], ywin, one, sr, ab);
                quad_chunk<4, IMADL>(cur6[4], ywin, one, sr, ab);
                quad_chunk<5, IMADL>(cur6[5], ywin, one, sr, ab);
                const unsigned SR = (sr[0] + sr[1]) + (sr[2] + sr[3]);
                const unsigned AB = (ab[0] + ab[1]) + (ab[2] + ab[3]);          // S1 | S3 << 16
                const unsigned AE = SR - (AB << 8);                             // S0 | S2 << 16
                const unsigned j = bcur * 32u + lane;
                const bool valid = (int)j < len;
                // thresholds are compared in the float domain (one FSETP per query); the order-preserving integer key is built only for
                // the rare vector that passes.  (float)(unsigned short) converts a 16-bit half of the register without a mask / shift.
                const float f0 = fmaf(stepv[0], (float)(unsigned short)(AE), basev[0]);
                const float f1 = fmaf(stepv[1], (float)(unsigned short)(AB), basev[1]);
                const float f2 = fmaf(stepv[2], (float)(unsigned short)(AE >> 16), basev[2]);
                const float f3 = fmaf(stepv[3], (float)(unsigned short)(AB >> 16), basev[3]);
                const bool p0 = valid && f0 >= tf0, p1 = valid && f1 >= tf1, p2 = valid && f2 >= tf2, p3 = valid && f3 >= tf3;
                if (__any_sync(0xffffffffu, p0 || p1 || p2 || p3)) {
                    const unsigned k0 = dph_fkey(f0), k1 = dph_fkey(f1), k2 = dph_fkey(f2), k3 = dph_fkey(f3);
                    warp_append_latch(p0 && k0 >= thr0, ((unsigned long long)k0 << 32) | (unsigned long long)(0xFFFFFFFFu - (gsv[0] + j)), sh->cbuf[0], &sh->cnt[0], &sh->full, lane);
                    warp_append_latch(p1 && k1 >= thr1, ((unsigned long long)k1 << 32) | (unsigned long long)(0xFFFFFFFFu - (gsv[1] + j)), sh->cbuf[1], &sh->cnt[1], &sh->full, lane);
                    warp_append_latch(p2 && k2 >= thr2, ((unsigned long long)k2 << 32) | (unsigned long long)(0xFFFFFFFFu - (gsv[2] + j)), sh->cbuf[2], &sh->cnt[2], &sh->full, lane);
                    warp_append_latch(p3 && k3 >= thr3, ((unsigned long long)k3 << 32) | (unsigned long long)(0xFFFFFFFFu - (gsv[3] + j)), sh->cbuf[3], &sh->cnt[3], &sh->full, lane);
                }
            }
            if (!more && !counted) { counted = true; if (lane == 0) atomicAdd(&sh->ndone, 1); }
            __syncthreads();
#pragma unroll 1
            for (int i = 0; i < nq; i++)
                if (sh->cnt[i] > a.keep) compact_buffer<QCAP>(sh->cbuf[i], &sh->cnt[i], &sh->thr[i], a.keep, a.gthr + sh->qs[i], &sh->sc);
            if (tid == 0) {
                for (int i = 0; i < nq; i++) { unsigned g = *((volatile unsigned*)(a.gthr + sh->qs[i])); if (g > sh->thr[i]) sh->thr[i] = g; }
                sh->ndone_snap = sh->ndone;
                sh->full = 0;
            }
            __syncthreads();
            thr0 = sh->thr[0]; thr1 = sh->thr[1]; thr2 = sh->thr[2]; thr3 = sh->thr[3];
            tf0 = thr_f(thr0); tf1 = thr_f(thr1); tf2 = thr_f(thr2); tf3 = thr_f(thr3);
            if (sh->ndone_snap == NW) break;
        }
        // ---- publish the candidate sets: the (up to) four region reservations go out together, one barrier ----
        if (tid < nq) sh->base[tid] = atomicAdd(a.cand_cnt + sh->qs[tid], sh->cnt[tid]);
        if (warp == 0) asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();
#pragma unroll 1
        for (int i = 0; i < nq; i++) {
            const long long q = sh->qs[i];
            const int cnt = sh->cnt[i];
            const long long off = a.cand_off[q], cap = a.cand_off[q + 1] - off;
            const int basep = sh->base[i];
            for (int c = tid; c < cnt; c += NT)
                if (basep + c < cap) a.cand[off + basep + c] = sh->cbuf[i][c];
        }
        const int un = sh->unit;
        __syncthreads();             // buffers, counters, sh->unit and the descriptor slots are reused by the next item
        if (un >= total_units) break;
        cur ^= 1;
    }
}

__global__ void smem_base_probe_kernel(unsigned* out) { *out = (unsigned)__cvta_generic_to_shared(dph_smem) & 0x00FFFFFFu; }

int dph_scan_setup_attrs() {
    static DphPerDeviceOnce once;
    if (!once.first()) return 0;
    {
        unsigned* d = nullptr; unsigned h = 0;
        DPH_CUDA(cudaMalloc((void**)&d, 4));
        smem_base_probe_kernel<<<1, 32, 1024>>>(d);
        DPH_CUDA(cudaMemcpy(&h, d, 4, cudaMemcpyDeviceToHost));
        cudaFree(d);
        DPH_CHECK(h == DPH_DYN_SMEM_BASE, "dynamic shared memory window does not start at 0x400 on this driver; rebuild with the probed value");
    }
    DPH_CUDA(cudaFuncSetAttribute(scan_kernel<DPH_SCAN_FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LUT_FAST + (int)sizeof(ScanShared)));
    DPH_CUDA(cudaFuncSetAttribute(scan_kernel<DPH_SCAN_EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LUT_EXACT + (int)sizeof(ScanShared)));
    DPH_CUDA(cudaFuncSetAttribute(scan_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LUT_FAST + (int)sizeof(PairShared)));
    DPH_CUDA(cudaFuncSetAttribute(scan_quad_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LUT_FAST + (int)sizeof(QuadShared)));
    DPH_CUDA(cudaFuncSetAttribute(scan_quad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LUT_FAST + (int)sizeof(QuadShared)));
    DPH_CUDA(cudaFuncSetAttribute(scan_quad_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LUT_FAST + (int)sizeof(QuadShared)));
    DPH_CUDA(cudaFuncSetAttribute(scan_quad_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LUT_FAST + (int)sizeof(QuadShared)));
    return 0;
}

int dph_launch_scan(dph_index* ix, int64_t n, int k, int keep, int mode, int grid, cudaStream_t st) {
    (void)k;
    if (n == 0) return 0;
    DPH_TRY(dph_scan_setup_attrs());
    ScanArgs a;
    a.codes = ix->codes; a.qpre = ix->wpre.as<long long>(); a.segs = ix->segs.as<DphSeg>(); a.nseg = ix->nseg.as<int>(); a.work = ix->work.as<DphWork>();
    a.lut_canon = ix->lut_canon.as<float>(); a.gthr = ix->gthr.as<unsigned>();
    a.cand = ix->cand.as<unsigned long long>(); a.cand_off = ix->cand_off.as<long long>(); a.cand_cnt = ix->cand_cnt.as<int>();
    a.n = n; a.nprobe = ix->nprobe; a.keep = keep;
    if (mode == DPH_SCAN_FAST)
        scan_kernel<DPH_SCAN_FAST><<<grid, NT, SMEM_LUT_FAST + sizeof(ScanShared), st>>>(a);
    else
        scan_kernel<DPH_SCAN_EXACT><<<grid, NT, SMEM_LUT_EXACT + sizeof(ScanShared), st>>>(a);
    DPH_CUDA(cudaGetLastError());
    return 0;
}

int dph_launch_scan_pair(dph_index* ix, int64_t n, int keep, int grid, cudaStream_t st, int group) {
    if (n == 0) return 0;
    DPH_TRY(dph_scan_setup_attrs());
    PairScanArgs a;
    a.codes = ix->codes; a.blk_off = (const long long*)ix->blk_off; a.list_len = ix->list_len; a.pl_cnt = ix->pl_cnt.as<int>();
    a.pl_off = ix->pl_off.as<int>(); a.units = ix->pl_units.as<unsigned long long>(); a.entries = ix->pl_entries.as<unsigned>();
    a.work = ix->pairwork.as<DphPairWork>(); a.next_unit = &ix->pairwork.as<DphPairWork>()->next_unit; a.lutq = ix->lutq.as<unsigned short>(); a.qparams = ix->qparams.as<float2>();
    a.cd = ix->cd.as<float>(); a.gdense = ix->gdense.as<unsigned>(); a.gthr = ix->gthr.as<unsigned>();
    a.cand = ix->cand.as<unsigned long long>(); a.cand_off = ix->cand_off.as<long long>(); a.cand_cnt = ix->cand_cnt.as<int>();
    a.list_lo = ix->list_lo; a.list_hi = ix->list_hi; a.nprobe = ix->nprobe; a.keep = keep;
    a.udesc = ix->pl_udesc.as<DphUnit>(); a.one = 1u;
    if (group == 4) {
        const size_t sm = SMEM_LUT_FAST + sizeof(QuadShared);
        switch (g_dph_tune[0]) {
            case 0: scan_quad_kernel<0><<<grid, NT, sm, st>>>(a); break;
            case 2: scan_quad_kernel<2><<<grid, NT, sm, st>>>(a); break;
            case 3: scan_quad_kernel<3><<<grid, NT, sm, st>>>(a); break;
            default: scan_quad_kernel<1><<<grid, NT, sm, st>>>(a); break;
        }
    }
    else scan_pair_kernel<<<grid, NT, SMEM_LUT_FAST + sizeof(PairShared), st>>>(a);
    DPH_CUDA(cudaGetLastError());
    return 0;
}

// =================================================================================================
// merge: one CTA per query.  FAST: prove the filter, re-score survivors in canonical order, sort.
// EXACT: candidates already carry canonical scores.  Output: D (score desc), I (labels), G (scan position).
// =================================================================================================
struct MergeArgs {
    const unsigned long long* cand; const long long* cand_off; const int* cand_cnt; const unsigned* gthr; const float* eps;
    const DphSeg* segs; const int* nseg; int nprobe; int k; int mode;
    const uint8_t* codes; const float* lut_canon; const long long* ids; const long long* list_start;
    float* D; long long* I; unsigned* G; int* flags; const int* only_flagged;
};

// last segment r with gstart[r] <= gidx; gstart = the query's segment starts, staged in shared memory by the caller (a binary search
// over the descriptors in global memory is a chain of ~8 dependent L2 round trips per survivor)
__device__ __forceinline__ int find_seg(const unsigned* gstart, int nsg, unsigned gidx) {
    int lo = 0, hi = nsg;
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (gstart[mid] <= gidx) lo = mid; else hi = mid; }
    return lo;
}

__global__ void __launch_bounds__(256) merge_kernel(MergeArgs a) {
    __shared__ SelectScratch sc;
    __shared__ unsigned long long surv[DPH_SURV_CAP];
    __shared__ unsigned sgs[DPH_MAX_NPROBE];
    __shared__ int scnt, sflag;
    const long long q = blockIdx.x;
    const int tid = threadIdx.x;
    if (a.only_flagged && a.only_flagged[q] == 0) return;
    const unsigned long long* E = a.cand + a.cand_off[q];
    long long cap = a.cand_off[q + 1] - a.cand_off[q];
    int n = a.cand_cnt[q];
    const bool overflow = n > cap;          // cannot happen with the plan's capacity bound; if it ever does the query goes to the exact kernel
    if (n > cap) n = (int)cap;
    const DphSeg* segs = a.segs + q * a.nprobe;
    const int nsg = a.nseg[q];
    const int k = a.k;
    if (tid == 0) { scnt = 0; sflag = 0; }
    for (int i = tid; i < nsg; i += blockDim.x) sgs[i] = segs[i].gstart;
    __syncthreads();
    auto get = [&](int i) { return E[i]; };
    unsigned long long pivot = 0ull;   // gather keys >= pivot
    int flag = overflow ? 1 : 0;
    if (a.mode == DPH_SCAN_FAST) {
        const float eps2 = 2.0f * a.eps[q];
        const unsigned gt = a.gthr[q];
        if (n > k) {
            unsigned long long pk = block_radix_select(get, n, k, &sc);
            const float Tk = dph_ckey_score(pk);
            if (gt != 0u && !(Tk - dph_fkey_inv(gt) > eps2)) flag = 1;
            pivot = (unsigned long long)dph_fkey(Tk - eps2) << 32;       // every entry with filter score >= Tk - 2eps
        } else if (gt != 0u) flag = 1;
    } else {
        if (n > DPH_SURV_CAP) pivot = block_radix_select(get, n, k, &sc);
    }
    for (int i = tid; i < n; i += blockDim.x) {
        unsigned long long e = E[i];
        if (e >= pivot) { int p = atomicAdd(&scnt, 1); if (p < DPH_SURV_CAP) surv[p] = e; else sflag = 1; }
    }
    __syncthreads();
    int ns = scnt < DPH_SURV_CAP ? scnt : DPH_SURV_CAP;
    if (sflag) flag = 1;
    if (a.mode == DPH_SCAN_FAST) {
        // exact re-scoring, one WARP per survivor: lane l fetches the code byte and the LUT entry of sub-quantizers l, l+32, l+64
        // (96 independent loads in flight instead of a 96-long chain of dependent L2 round trips), then the lanes' values are added
        // in canonical m-ascending order (bit-exact with the oracle: dis = dis0; for m: dis += LUT[m][c[m]]).
        const float* lutc = a.lut_canon + (size_t)q * DPH_LUT_CANON_FLOATS;
        const int lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
        for (int i = warp; i < ns; i += nwarps) {
            const unsigned gidx = dph_ckey_gidx(surv[i]);
            const DphSeg s = segs[find_seg(sgs, nsg, gidx)];
            const unsigned j = gidx - s.gstart;
            const uint8_t* blk = a.codes + (s.blk + (long long)(j >> 5)) * DPH_BLK_BYTES;
            const int ln = (int)(j & 31u);
            float v[3];
#pragma unroll
            for (int h = 0; h < 3; h++) {
                const int m = h * 32 + lane;
                v[h] = __ldg(lutc + DPH_LUTC_IDX(m, (int)blk[dph_blk_addr(ln, m)]));
            }
            float dis = s.dis0;
#pragma unroll
            for (int h = 0; h < 3; h++)
#pragma unroll
                for (int l2 = 0; l2 < 32; l2++) dis += __shfl_sync(0xffffffffu, v[h], l2);
            __syncwarp();
            if (lane == 0) surv[i] = dph_ckey(dis, gidx);
        }
    }
    const int p2 = dph_next_pow2(ns > 1 ? ns : 1);
    for (int i = ns + tid; i < p2; i += blockDim.x) surv[i] = 0ull;
    __syncthreads();
    block_bitonic_sort_desc(surv, p2);
    for (int i = tid; i < k; i += blockDim.x) {
        float d = DPH_NEUTRAL; long long id = -1; unsigned gi = 0xFFFFFFFFu;
        if (i < ns) {
            const unsigned long long e = surv[i];
            gi = dph_ckey_gidx(e); d = dph_ckey_score(e);
            const DphSeg s = segs[find_seg(sgs, nsg, gi)];
            const unsigned j = gi - s.gstart;
            id = a.ids ? a.ids[s.blk * 32 + j] : a.list_start[s.list] + (long long)j;
        }
        a.D[q * k + i] = d; a.I[q * k + i] = id; a.G[q * k + i] = gi;
    }
    if (a.mode == DPH_SCAN_FAST && tid == 0) a.flags[q] = flag;
}

int dph_launch_merge(dph_index* ix, int64_t n, int k, int mode, const int32_t* only_flagged, float* D, int64_t* I, uint32_t* G,
                     cudaStream_t st) {
    if (n == 0) return 0;
    MergeArgs a;
    a.cand = ix->cand.as<unsigned long long>(); a.cand_off = ix->cand_off.as<long long>(); a.cand_cnt = ix->cand_cnt.as<int>();
    a.gthr = ix->gthr.as<unsigned>(); a.eps = ix->eps.as<float>(); a.segs = ix->segs.as<DphSeg>(); a.nseg = ix->nseg.as<int>(); a.nprobe = ix->nprobe; a.k = k; a.mode = mode;
    a.codes = ix->codes; a.lut_canon = ix->lut_canon.as<float>(); a.ids = (const long long*)ix->ids; a.list_start = (const long long*)ix->list_start;
    a.D = D; a.I = (long long*)I; a.G = G; a.flags = ix->flags.as<int>(); a.only_flagged = only_flagged;
    merge_kernel<<<(unsigned)n, 256, 0, st>>>(a);
    DPH_CUDA(cudaGetLastError());
    return 0;
}

// =================================================================================================
// pack / merge of the per-shard partial top-k for ONE all-gather:  P[q][i] = { ckey(score, scan position), label }.
// =================================================================================================
__global__ void pack_topk_kernel(const float* D, const long long* I, const unsigned* G, long long total, longlong2* P) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    longlong2 v;
    v.x = I[i] >= 0 ? (long long)dph_ckey(D[i], G[i]) : 0ll;
    v.y = I[i];
    P[i] = v;
}
__global__ void __launch_bounds__(256) merge_packed_kernel(const longlong2* Pg, int nshards, long long n, int k, float* D, long long* I) {
    extern __shared__ unsigned long long mp[];      // [p2] keys, then [p2] labels
    const long long q = blockIdx.x;
    const int tot = nshards * k, p2 = dph_next_pow2(tot);
    unsigned long long* keys = mp;
    for (int i = threadIdx.x; i < p2; i += blockDim.x) {
        unsigned long long key = 0ull;
        if (i < tot) key = (unsigned long long)Pg[((long long)(i / k) * n + q) * k + (i % k)].x;
        keys[i] = key;
    }
    __syncthreads();
    block_bitonic_sort_desc(keys, p2);
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        float d = DPH_NEUTRAL; long long id = -1;
        const unsigned long long key = keys[i];
        if (key != 0ull) {
            d = dph_ckey_score(key);
            for (int j = 0; j < tot && id < 0; j++) {          // scan positions are unique per query: find the owner (<= nshards*k entries)
                const longlong2 e = Pg[((long long)(j / k) * n + q) * k + (j % k)];
                if ((unsigned long long)e.x == key) id = e.y;
            }
        }
        D[q * k + i] = d; I[q * k + i] = id;
    }
}
DPH_API int dph_pack_topk(const float* D, const int64_t* I, const uint32_t* G, int64_t n, int k, int64_t* P, void* cuda_stream) {
    const long long total = (long long)n * k;
    if (total == 0) return 0;
    pack_topk_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)cuda_stream>>>(D, (const long long*)I, G, total, (longlong2*)P);
    DPH_CUDA(cudaGetLastError());
    return 0;
}
DPH_API int dph_merge_shards_packed(const int64_t* Pg, int nshards, int64_t n, int k, float* D, int64_t* I, void* cuda_stream) {
    if (n == 0) return 0;
    DPH_CHECK(nshards >= 1 && k >= 1 && (long long)nshards * k <= 8192, "merge_shards: nshards*k must be <= 8192");
    int p2 = 1; while (p2 < nshards * k) p2 <<= 1;
    static DphPerDeviceOnce once;
    if (once.first()) { DPH_CUDA(cudaFuncSetAttribute(merge_packed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 8)); }
    merge_packed_kernel<<<(unsigned)n, 256, p2 * 8, (cudaStream_t)cuda_stream>>>((const longlong2*)Pg, nshards, n, k, D, (long long*)I);
    DPH_CUDA(cudaGetLastError());
    return 0;
}

// =================================================================================================
// merge_shards: all-gathered per-shard top-k -> global top-k, order (score desc, scan position asc).
// =================================================================================================
__global__ void __launch_bounds__(256) merge_shards_kernel(const float* Dg, const long long* Ig, const unsigned* Gg, int nshards, long long n,
                                                            int k, float* D, long long* I) {
    extern __shared__ unsigned long long ms[];     // keys [p2] then payload index is recovered from a parallel array
    const long long q = blockIdx.x;
    const int tid = threadIdx.x;
    const int tot = nshards * k;
    const int p2 = dph_next_pow2(tot);
    unsigned long long* keys = ms;
    // key = (fkey(score), ~gidx); payload looked up afterwards by matching (shard, slot) stored in a side table:
    // gidx is unique per real entry, so re-find the source by scanning the <= nshards*k inputs (tiny).
    for (int i = tid; i < p2; i += blockDim.x) {
        unsigned long long key = 0ull;
        if (i < tot) {
            int s = i / k, r = i % k;
            long long id = Ig[((long long)s * n + q) * k + r];
            if (id >= 0) key = dph_ckey(Dg[((long long)s * n + q) * k + r], Gg[((long long)s * n + q) * k + r]);
        }
        keys[i] = key;
    }
    __syncthreads();
    block_bitonic_sort_desc(keys, p2);
    for (int i = tid; i < k; i += blockDim.x) {
        float d = DPH_NEUTRAL; long long id = -1;
        unsigned long long key = keys[i];
        if (key != 0ull) {
            unsigned gi = dph_ckey_gidx(key);
            d = dph_ckey_score(key);
            for (int s = 0; s < nshards && id < 0; s++)
                for (int r = 0; r < k; r++) {
                    long long o = ((long long)s * n + q) * k + r;
                    if (Gg[o] == gi && Ig[o] >= 0) { id = Ig[o]; break; }
                }
        }
        D[q * k + i] = d; I[q * k + i] = id;
    }
}

DPH_API int dph_merge_shards(const float* Dg, const int64_t* Ig, const uint32_t* Gg, int nshards, int64_t n, int k, float* D, int64_t* I,
                                void* cuda_stream) {
    if (n == 0) return 0;
    DPH_CHECK(nshards >= 1 && k >= 1 && (long long)nshards * k <= 8192, "merge_shards: nshards*k must be <= 8192");
    int p2 = 1; while (p2 < nshards * k) p2 <<= 1;
    static DphPerDeviceOnce once;
    if (once.first()) { DPH_CUDA(cudaFuncSetAttribute(merge_shards_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 8)); }
    merge_shards_kernel<<<(unsigned)n, 256, p2 * 8, (cudaStream_t)cuda_stream>>>(Dg, (const long long*)Ig, Gg, nshards, n, k, D, (long long*)I);
    DPH_CUDA(cudaGetLastError());
    return 0;
}
