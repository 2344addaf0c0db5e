// umma.cuh -- thin PTX wrappers for the sm_100a tensor-core path (mbarrier, TMA tensor loads, tcgen05.mma / commit / ld,
// shared-memory matrix descriptors) shared by gemm_tf32.cu and attention_tc.cu, plus the host-side tensor-map encoder.
#pragma once
#include "common.cuh"
#include <cuda.h>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap* map, int c0, int c1, unsigned bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
// multicast variant: the box lands at the same CTA-relative offset (and signals the same barrier offset) in every CTA of cta_mask
__device__ __forceinline__ void tma_load_2d_mc(unsigned dst, const CUtensorMap* map, int c0, int c1, unsigned bar, unsigned short cta_mask) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], [%4], %5;"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma_tf32(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc, unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// arrive on the barrier at this CTA-relative offset in every CTA of cta_mask once the MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_mc(unsigned bar, unsigned short cta_mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(cta_mask) : "memory");
}
// K-major, SWIZZLE_128B operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart (cute::UMMA::SmemDescriptor, sm_100 version 1).
__device__ __forceinline__ unsigned long long make_sw128_desc(unsigned smem_addr) {
    unsigned long long d = 0;
    d |= (unsigned long long)((smem_addr & 0x3FFFF) >> 4);        // start address, bits [0,14)
    d |= (unsigned long long)1 << 16;                            // leading byte offset (ignored for swizzled K-major), bits [16,30)
    d |= (unsigned long long)(1024 >> 4) << 32;                  // stride byte offset, bits [32,46)
    d |= (unsigned long long)1 << 46;                            // descriptor version (Blackwell)
    d |= (unsigned long long)2 << 61;                            // layout type SWIZZLE_128B
    return d;
}

// 32 consecutive fp32 accumulator columns of this thread's TMEM lane (shape 32x32b: warp w of a warpgroup owns lanes 32w..32w+31)
__device__ __forceinline__ void tmem_ld32(unsigned taddr, unsigned (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- host: cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda) ----
typedef CUresult (*dph_PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int dph_tensormap_encoder(dph_PFN_encodeTiled* out);      // gemm_tf32.cu
// [rows, cols] fp32 row-major (row stride ld floats) -> map with a [32 floats x box_rows] box, 128-byte swizzle, zero fill out of bounds
int dph_make_map_f32(CUtensorMap* map, const float* ptr, long long rows, long long cols, long long ld, int box_rows);

// ---- bf16 operands (gemm_bf16x3.cu) ----
__device__ __forceinline__ void umma_bf16(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc, unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major, SWIZZLE_64B operand tile: rows of 64 bytes (32 bf16), 8-row groups 512 bytes apart.
__device__ __forceinline__ unsigned long long make_sw64_desc(unsigned smem_addr) {
    unsigned long long d = 0;
    d |= (unsigned long long)((smem_addr & 0x3FFFF) >> 4);        // start address, bits [0,14)
    d |= (unsigned long long)1 << 16;                            // leading byte offset (ignored for swizzled K-major), bits [16,30)
    d |= (unsigned long long)(512 >> 4) << 32;                   // stride byte offset, bits [32,46)
    d |= (unsigned long long)1 << 46;                            // descriptor version (Blackwell)
    d |= (unsigned long long)4 << 61;                            // layout type SWIZZLE_64B
    return d;
}
// [rows, cols] bf16 row-major (row stride ld elements) -> map with a [32 bf16 x box_rows] box, 64-byte swizzle, zero fill out of bounds
int dph_make_map_bf16(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld, int box_rows);
