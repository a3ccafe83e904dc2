// Thin inline-PTX layer over the Blackwell (sm_100a) tensor-core path used by the dense
// projection and LSTM kernels: tcgen05.mma with shared-memory operand descriptors,
// accumulators in TMEM, mbarrier completion, tcgen05.ld for the epilogue.
//
// Operand layout used everywhere in this library: K-major, NO swizzle ("interleave"):
//   an operand tile of R rows x K columns (bf16) is stored as 8-row x 8-column core matrices
//   of 128 contiguous bytes (row r of a core matrix at +16*r bytes);
//     byte offset(r, k) = (r / 8) * SBO + (k / 8) * LBO + (r % 8) * 16 + (k % 8) * 2
//   with LBO = 128 (core matrices adjacent along K are contiguous) and SBO = (K / 8) * 128.
// Bit layouts of the descriptors follow cute/arch/mma_sm100_desc.hpp (CUTLASS 3.9+):
//   smem descriptor: [0,14) addr>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version=1, [61,64) layout=0
//   instr descriptor (kind::f16): [4,6) D fmt (1=f32), [7,10) A fmt (1=bf16), [10,13) B fmt,
//                    bit 15/16 A/B major (0 = K-major), [17,23) N>>3, [24,29) M>>4
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace mlvae {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one lane of a fully active warp (use under a warp-uniform condition: the compiler then emits the
// tcgen05 instruction once instead of a serialising loop over the active lanes)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred)::"memory");
    return pred != 0;
}

// ---- mbarrier --------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// generic-proxy writes (st.shared / cp.async) -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- cp.async (16 bytes, zero-fill when src_bytes == 0) ----------------------------------
__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src, uint32_t src_bytes = 16) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "r"(src_bytes) : "memory");
}
// arrive on `bar` (without incrementing its pending count) once all cp.async issued so far by this thread have landed
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- TMEM ------------------------------------------------------------------------------
// one full warp; ncols power of two in [32, 512]; the TMEM base address lands in *dst_smem
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- descriptors ------------------------------------------------------------------------
__device__ __forceinline__ uint64_t smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;          // descriptor version 1 (Blackwell)
    return d;                        // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}
// K-major SWIZZLE_128B operand tile (rows of exactly 64 bf16 = 128 bytes; 8-row atoms of 1024 bytes; 16-byte chunk c of
// row r stored at chunk position c ^ (r & 7)): tile base 1024-byte aligned, SBO = 1024, LBO unused, layout_type = 2.
// A K = 16 step inside the 128-byte row advances the start address by 32 bytes.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t sw128_off(int r, int c16) {        // byte offset of 16-byte chunk c16 (0..7) of row r
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c16 ^ (r & 7)) << 4));
}
__host__ __device__ constexpr uint32_t idesc_bf16_f32(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- MMA --------------------------------------------------------------------------------
// D[tmem] (+)= A[smem desc] * B[smem desc]^T ; issued by ONE thread on behalf of the CTA
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc),
        "r"((uint32_t)accumulate)
        : "memory");
}
// same with the A operand resident in TMEM (lane = row m, 32-bit column j holds K elements 2j, 2j+1)
__device__ __forceinline__ void mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc),
        "r"((uint32_t)accumulate)
        : "memory");
}
// registers -> TMEM: warp w writes lanes [32*(w%4), +32), 8 consecutive 32-bit columns per thread
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// arrive on `bar` once every MMA issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- TMEM -> registers: warp w (of a 4-warp group) reads lanes [32*(w%4), +32), 16 consecutive columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&v)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
                 : "r"(taddr)
                 : "memory");
}
// N consecutive columns (N in {4, 8, 16}) of this warp's 32 lanes -> v[0..N)
template <int N> __device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t (&v)[N]) {
    if constexpr (N == 4) tmem_ld4(taddr, v);
    else if constexpr (N == 8) tmem_ld8(taddr, v);
    else tmem_ld16(taddr, v);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 consecutive columns -> v[0..32)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}

// ---- TMA (cp.async.bulk.tensor) --------------------------------------------------------------
// arm `bar` with the number of bytes the TMA loads of this stage will deposit (+ one arrival)
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global (tensor map, 3-D coordinates c0 = innermost) -> shared, completion reported to `bar` in bytes
__device__ __forceinline__ void tma_load_3d(void *dst_smem, const void *tmap, uint64_t *bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(dst_smem)),
        "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// shared -> global (tensor map); completion tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void tma_store_3d(const void *tmap, const void *src_smem, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(tmap), "r"(smem_u32(src_smem)), "r"(c0),
                 "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until at most N of this thread's bulk groups still READ their shared-memory source
template <int N> __device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void tma_store_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const void *tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
// MN-major SWIZZLE_128B operand tile as TMA deposits it: blocks of 64 MN elements (128 bytes) x `krows` K rows (krows * 128
// bytes per block, blocks `lbo_bytes` apart); 8-row swizzle atoms of 1024 bytes (SBO).  A K = 16 step advances the start
// address by 2048 bytes.  (cute/atom/mma_traits_sm100.hpp: ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units.)
__device__ __forceinline__ uint64_t smem_desc_sw128_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// instruction descriptor with operand major-ness: bit 15 / 16 = A / B is MN-major
__host__ __device__ constexpr uint32_t idesc_bf16_f32_major(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(a_mn & 1) << 15) | ((uint32_t)(b_mn & 1) << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}

// byte offset of element (r, k) inside a no-swizzle K-major tile with `kdim` columns
__host__ __device__ constexpr uint32_t kmajor_off(int r, int k, int kdim) {
    return (uint32_t)((r >> 3) * (kdim >> 3) * 128 + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2);
}

}  // namespace tc
}  // namespace mlvae
