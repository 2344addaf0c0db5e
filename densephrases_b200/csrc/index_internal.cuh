// index_internal.cuh -- the dph_index handle and the internal kernel-launch prototypes.
#pragma once
#include "common.cuh"
#include "../../include/dph_b200.h"
#include <vector>

#define DPH_SCAN_THREADS 512
#define DPH_SCAN_WARPS (DPH_SCAN_THREADS / 32)
#define DPH_CAND_CAP 3072          // shared-memory candidate buffer (u64 keys) per scan CTA
#define DPH_PAIR_CAP 1280          // pair mode: one buffer per query of the pair
#define DPH_QUAD_KEEP_MAX 256       // quad mode: QCAP (768) - scan threads (512), see scan.cu

#define DPH_KEEP_SLACK 32          // fast mode keeps k + slack candidates per CTA
#define DPH_MAX_K 1024
#define DPH_PROF_RING 64
#define DPH_MAX_NPROBE 1024
#define DPH_SURV_CAP 2048          // merge kernel: survivors re-scored exactly per query
#define DPH_LUT_SCAN_FLOATS (3 * 256 * 64)   // per query: 3 segments x 256 codes x (32 + 31 dup + 1 pad)
#define DPH_LUT_CANON_FLOATS (96 * 256)
// The canonical fp32 table of a query, LUT[m][code] = <xr[8m..8m+8), pq[m][code]>, is stored segment-major, code-major,
// sub-quantizer-minor: [3 segments of 32 sub-quantizers][256 codes][32] -- i.e. already in the row order the conflict-free scan
// layout needs (scan.cu), so ONE 96 KB table per query serves the one-query scan (staged with its wrap copies), the quantised
// pair / quad tables, the exact kernel and the exact re-scoring in the merge.
#define DPH_LUTC_IDX(m, code) ((((m) >> 5) << 13) + ((code) << 5) + ((m) & 31))
#define DPH_SEG_SMEM 256            // segment descriptors of one query kept in shared memory by the scan kernel
#define DPH_L2_PREFETCH_ROUNDS 4    // scan kernel: bulk L2 prefetch distance, in rounds (one 3 KB block per warp per round)

struct DevBuf {            // grow-only device buffer
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes);
    void release();
    template <class T> T* as() const { return (T*)p; }
};

struct dph_index {
    int device = 0;
    int d = DPH_D, M = DPH_M;
    int64_t nlist = 0;
    int nprobe = 256;
    int scan_mode = DPH_SCAN_FAST;
    cudaStream_t stream = 0;
    int num_sms = 148;

    // model
    float* A = nullptr;         // [d,d]
    float* C = nullptr;         // [nlist,d]
    float* pq = nullptr;        // [M,256,dsub]
    // lists
    int64_t list_lo = 0, list_hi = 0;   // shard range
    int64_t ntotal = 0, ntotal_local = 0, nblocks_local = 0;
    int32_t* list_len = nullptr;        // [nlist]   (all lists)
    int64_t* list_start = nullptr;      // [nlist+1] global list-major row of each list's first vector
    int64_t* blk_off = nullptr;         // [nlist]   first code block in `codes` (-1 when not in shard)
    uint8_t* codes = nullptr;           // [nblocks_local * 3072] interleaved blocks
    int64_t* ids = nullptr;             // [nblocks_local * 32] labels (or nullptr: sequential)
    // direct map for explicit ids (sorted labels -> local padded row); built at set_lists
    int64_t* dm_ids = nullptr;
    int64_t* dm_rows = nullptr;
    int64_t dm_n = 0;
    std::vector<int64_t> h_list_len;    // host copies
    std::vector<int64_t> h_list_start;
    int64_t bytes = 0;

    // per-batch workspace
    DevBuf xdev, xr, S, key, cd, lut_canon, lutmax, segs, wpre, qinfo, cand, cand_off, cand_cnt, gthr, flags,
        work, Dp, Ip, Gp, Dh, Ih, eps, nseg,
        lutmin, lutmaxv, lutq, qparams, gdense, pl_cnt, pl_fill, pl_off, pl_blockpre, pl_entries, pl_unitpre, pl_units, pl_udesc, pairwork,
        csplit, xsplit, candkeys, cflags, selkeys, recbuf,
        rb_ids, rb_out, rb_found, ws_q, ws_id, ws_out, ws_xq;        // reconstruct_batch / window_scores staging (host-buffer calls)
    int64_t csplit_lo = -1, csplit_nl = -1;
    int coarse_tc = 1;                 // tensor-core coarse quantizer with exact re-rank (0: always the SIMT sequential-k GEMM)
    int64_t last_n = 0;
    int last_group = 1;                // queries per gather used by the last search (1, 2 or 4)
    int64_t last_coarse_n = -1;
    bool profile = false;              // CUDA events around the scan kernel of the last search chunk
    cudaEvent_t ev0[DPH_PROF_RING] = {}, ev1[DPH_PROF_RING] = {};
    int64_t prof_n = 0;
};

// process-wide variant selection (dph_set_tuning, measurement hook): [0] quad-scan IMAD level, [1] SGEMM tile
extern int g_dph_tune[8];
// ---- prep.cu ----
int dph_launch_sgemm_nt_seq(const float* X, int64_t n, const float* W, int64_t m, int K, float* out, cudaStream_t st);
int dph_launch_coarse_select(const float* S, int64_t n, int64_t nlist, int nprobe, int32_t* key, float* cd, cudaStream_t st,
                             unsigned long long* keys64 = nullptr, unsigned list_base = 0, const int* only_rows = nullptr, int64_t ld = 0,
                             DevBuf* tmp = nullptr);      // tmp: scratch for the chunked selection of long rows (nullptr: one CTA per row)
int dph_coarse_tc(dph_index* ix, int64_t n, int64_t lo, int64_t nl, int nprobe, unsigned long long* keys64, int32_t* key, float* cd, cudaStream_t st);
int dph_launch_coarse_merge(const unsigned long long* keys, int W, int64_t n, int nprobe, int32_t* key, float* cd, cudaStream_t st,
                            unsigned long long* keys64 = nullptr);
int dph_launch_lut(const float* xr, int64_t n, const float* pq, float* lut_canon, float* lutmax, float* lutmin, float* lutmaxv,
                   void* lutq, float2* qparams, cudaStream_t st, int group);
// group: queries per gather of the scan -- 1 (fp32 LUT, one query), 2 (pair-packed u16 LUTs), 4 (quad-packed u8 LUTs)
int dph_launch_plan(dph_index* ix, int64_t n, int k, int keep, int grid, const int32_t* only_flagged, cudaStream_t st, int group);
int dph_launch_scan_pair(dph_index* ix, int64_t n, int keep, int grid, cudaStream_t st, int group);
// ---- scan.cu ----
int dph_launch_scan(dph_index* ix, int64_t n, int k, int keep, int mode, int grid, cudaStream_t st);
int dph_launch_merge(dph_index* ix, int64_t n, int k, int mode, const int32_t* only_flagged, float* D, int64_t* I,
                     uint32_t* G, cudaStream_t st);
int dph_scan_setup_attrs();
