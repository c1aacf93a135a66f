// Blackwell (sm_100a) primitives shared by the tensor-core kernels: mbarrier, TMA, tcgen05 (MMA / TMEM), clusters.
// Thin inline-PTX wrappers only; no policy.
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace ef_tc {

constexpr int BLOCK_M = 128;                        // crops per tile (UMMA M, cta_group::1)
constexpr int BLOCK_K = 128;                        // bytes of K per stage = one 128-byte swizzle row
constexpr int UMMA_K = 32;                          // K per tcgen05.mma for 8-bit operands
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K;    // 16 KB
constexpr int kMaxStages = 8;
constexpr int kSmemLimit = 227 * 1024;
constexpr unsigned long long kTimeoutNs = 2000000000ull;

struct Shared {
  unsigned long long full_bar[kMaxStages];
  unsigned long long empty_bar[kMaxStages];
  unsigned long long tmem_full_bar;
  unsigned long long tmem_empty_bar;
  uint32_t tmem_base;
  int failed;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ unsigned long long globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: false on timeout or when another role already failed.
__device__ __forceinline__ bool mbar_wait(unsigned long long* bar, uint32_t parity, volatile int* failed) {
  // fast path: plain try_wait spins (each try_wait already suspends the thread for a hardware-defined interval);
  // the (slow) global timer and the shared failure flag are only consulted every 1024 unsuccessful polls.
  unsigned long long t0 = 0;
  for (unsigned int polls = 1;; ++polls) {
    if (mbar_try_wait(bar, parity)) return true;
    if ((polls & 1023u) == 0) {
      if (*failed) return false;
      const unsigned long long now = globaltimer();
      if (t0 == 0) t0 = now;
      if (now - t0 > kTimeoutNs) {
        *failed = 1;
        return false;
      }
    }
  }
}

// Cluster-scope variants for barriers that are arrived on by OTHER CTAs of the cluster after they wrote this CTA's shared
// memory (st.shared::cluster): the wait acquires at cluster scope, the remote arrive releases at cluster scope.
__device__ __forceinline__ bool mbar_try_wait_cluster(unsigned long long* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_wait_cluster(unsigned long long* bar, uint32_t parity, volatile int* failed) {
  unsigned long long t0 = 0;
  for (unsigned int polls = 1;; ++polls) {
    if (mbar_try_wait_cluster(bar, parity)) return true;
    if ((polls & 1023u) == 0) {
      if (*failed) return false;
      const unsigned long long now = globaltimer();
      if (t0 == 0) t0 = now;
      if (now - t0 > kTimeoutNs) {
        *failed = 1;
        return false;
      }
    }
  }
}
// arrive on a barrier in another CTA's shared memory (address from map_to_cta)
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

// Remote arrive without a memory fence: for hand-backs that release nothing (the arriving thread only READ the buffer it
// returns).  The .release.cluster form above costs a MEMBAR.ALL.GPU per arrive -- microseconds under a busy TMA stream.
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

// Asynchronous remote stores (STAS): the data and its completion travel together -- the store itself signals
// `bytes` on the destination CTA's mbarrier when it has landed, so the producer needs neither a fence nor an arrive.
__device__ __forceinline__ void st_async_u64(uint32_t cluster_addr, unsigned long long v, uint32_t cluster_mbar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];"
               ::"r"(cluster_addr), "l"(v), "r"(cluster_mbar) : "memory");
}
__device__ __forceinline__ void st_async_v2_u64(uint32_t cluster_addr, unsigned long long a, unsigned long long b,
                                                uint32_t cluster_mbar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];"
               ::"r"(cluster_addr), "l"(a), "l"(b), "r"(cluster_mbar) : "memory");
}

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, unsigned long long* bar, int c_inner,
                                            int c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer)
      : "memory");
}

// TMA prefetch of a 2-D box into L2 (no shared memory, no barrier): hides the HBM latency of a later tma_load_2d
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c_inner, int c_outer) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(map), "r"(c_inner), "r"(c_outer)
               : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// D[tmem] (+)= A[smem] * B[smem], u8 x s8 -> s32
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Shared-memory matrix descriptor: K-major operand, 128-byte swizzle, rows of 128 bytes, 8-row atoms 1024 bytes apart.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);   // start address,  bits [0,14)
  d |= (uint64_t)1 << 16;                        // leading byte offset (unused with swizzle), bits [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset = 8 rows * 128 B, bits [32,46)
  d |= (uint64_t)1 << 46;                        // descriptor version for sm_100
  d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
  return d;
}

// Instruction descriptor for kind::i8: D = s32, A = u8 (K-major), B = s8 (K-major), M = 128, N = n.
__host__ __device__ constexpr uint32_t umma_idesc_i8(int n) {
  return (2u << 4)                 // c_format = S32
         | (0u << 7)               // a_format = unsigned 8-bit
         | (1u << 10)              // b_format = signed 8-bit
         | (0u << 15) | (0u << 16) // both K-major
         | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// D[tmem] (+)= A[smem] * B[smem], f16 x f16 -> f32
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Instruction descriptor for kind::f16: D = f32, A = B = f16 (K-major), M = 128, N = n.
__host__ __device__ constexpr uint32_t umma_idesc_f16(int n) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
}

// Shared-memory matrix descriptor: K-major operand WITHOUT swizzle.  Canonical layout (cute: ((8,n),2):((1,SBO),LBO) in
// 16-byte units): core matrix = 8 rows x 16 bytes stored contiguously (128 B); lbo = byte distance between the two core
// matrices adjacent in K, sbo = byte distance between 8-row groups.
__device__ __forceinline__ uint64_t umma_desc_nosw(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;                        // descriptor version for sm_100
  return d;                                      // layout type 0 = no swizzle
}

// 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------------------------------------- clusters
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ int4 ld_cluster_v4(uint32_t addr) {
  int4 v;
  asm volatile("ld.shared::cluster.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ int ld_cluster_s32(uint32_t addr) {
  int v;
  asm volatile("ld.shared::cluster.s32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_cluster_u64(uint32_t addr) {
  unsigned long long v;
  asm volatile("ld.shared::cluster.u64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void st_cluster_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_u64(uint32_t addr, unsigned long long v) {
  asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}

__device__ __forceinline__ void st_cluster_v2_u64(uint32_t addr, unsigned long long a, unsigned long long b) {
  asm volatile("st.shared::cluster.v2.u64 [%0], {%1, %2};" ::"r"(addr), "l"(a), "l"(b) : "memory");
}

// K-major operand image with the widest swizzle that fits the row: rows of row_bytes in {32, 64, 128, 256} bytes
// (256 = two 128-byte swizzle atoms side by side in K).  Byte offset of 16-byte chunk `chunk` of row `row` in an image
// of `rows` rows whose base is 1024-byte aligned (Swizzle<1|2|3,4,3>: chunk index XOR row bits).
__host__ __device__ __forceinline__ uint32_t swz_chunk_offset(int row, int chunk, int row_bytes, int rows) {
  const int swb = row_bytes < 128 ? row_bytes : 128;
  const int cps = swb >> 4;                                   // chunks per swizzle row: 2, 4, 8
  const int shift = cps == 8 ? 0 : (cps == 4 ? 1 : 2);
  const int cps_log2 = cps == 8 ? 3 : (cps == 4 ? 2 : 1);
  const int katom = chunk >> cps_log2, cin = chunk & (cps - 1);
  return (uint32_t)(katom * rows * swb + row * swb + ((cin ^ ((row >> shift) & (cps - 1))) << 4));
}
// Matching shared-memory descriptor for k-step `ks` (32 bytes of K) of such an image.
__device__ __forceinline__ uint64_t umma_desc_swz(uint32_t base_addr, int ks, int row_bytes, int rows) {
  const int swb = row_bytes < 128 ? row_bytes : 128;
  const int per_atom = swb >> 5;                              // k-steps per swizzle atom: 1, 2, 4
  const int pa_log2 = per_atom == 4 ? 2 : (per_atom == 2 ? 1 : 0);
  const int katom = ks >> pa_log2, kin = ks & (per_atom - 1);
  const uint32_t addr = base_addr + (uint32_t)(katom * rows * swb + kin * 32);
  const uint64_t layout = swb == 128 ? 2ull : (swb == 64 ? 4ull : 6ull);
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;                                     // leading byte offset: unused inside one swizzle atom
  d |= (uint64_t)((8 * swb) >> 4) << 32;                      // stride byte offset: 8 rows
  d |= (uint64_t)1 << 46;
  d |= layout << 61;
  return d;
}

// --------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn();
// 2-D byte tensor [rows][pitch] with box [box_rows][128 bytes], 128-byte swizzle, zero fill out of bounds.
bool make_map(CUtensorMap* map, const void* base, uint64_t inner, uint64_t rows, uint64_t pitch, uint32_t box_rows);

}  // namespace ef_tc
