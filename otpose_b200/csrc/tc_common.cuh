// Blackwell (sm_100a) tensor-core plumbing: tcgen05 / TMEM / mbarrier PTX wrappers
// and the shared-memory operand layout used by every tcgen05 kernel in this library.
//
// Operand layout ("core-matrix interleaved", UMMA SWIZZLE_NONE): a bf16 tile with
// R rows and K columns (K contiguous within 16-byte chunks of 8 elements) is stored as
//
//     addr(r, k) = (r/8)*RS + (k/8)*CS + (r%8)*16 + (k%8)*2        [bytes]
//
// i.e. 8x8 "core matrices" of 128 contiguous bytes.  The same bytes serve as
//   * a K-major operand   (rows = M or N index, k = K index):  SBO = RS, LBO = CS,
//     K-step of 16 elements = start + 2*CS;
//   * an MN-major operand (k index of the layout = M/N index, rows = K index):
//     SBO = CS, LBO = RS, K-step of 16 = start + 2*RS
// so a tile written once (one thread per row, 16-byte stores: bank-conflict free)
// can feed a GEMM along either of its axes.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace otp {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor (SWIZZLE_NONE, version 1).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;
}

// Instruction descriptor, kind::f16: 16-bit x 16-bit -> fp32, M = 128.
// fmt: 1 = bf16 operands, 0 = fp16 operands (UMMA F16F32Format).
__host__ __device__ constexpr uint32_t make_idesc_16(int n, bool a_mn_major, bool b_mn_major, uint32_t fmt,
                                                     int m = 128) {
  return (1u << 4)                       // D format  = F32
         | (fmt << 7)                    // A format
         | (fmt << 10)                   // B format
         | ((a_mn_major ? 1u : 0u) << 15)
         | ((b_mn_major ? 1u : 0u) << 16)
         | ((uint32_t)(n >> 3) << 17)
         | ((uint32_t)(m >> 4) << 24);
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// Bounded wait: a protocol bug traps (launch error) after ~2 s instead of hanging the GPU.
// try_wait carries a suspend-time hint, so a waiting warp sleeps in hardware until the phase
// completes instead of competing for issue slots with the warps it is waiting for.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(addr), "r"(parity)
      : "memory");
  if (done) return;
  const long long t0 = clock64();
  for (;;) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity), "r"(100000u)
        : "memory");
    if (done) return;
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}

// ---------------------------------------------------------------- TMA (non-tensor bulk copy)
// One thread: arm the barrier with the byte count, then copy a contiguous global block
// (16-byte aligned, multiple of 16 bytes) into shared memory; the copy's completion
// (complete_tx) releases the barrier phase.
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------- TMA (tensor-map tiles)
// Host: descriptor of a row-major 2-D fp32 tensor (rows x cols, cols contiguous, row pitch = cols * 4 bytes,
// which has to be a multiple of 16) accessed in (box_rows x box_cols) tiles, no swizzle: the tile lands in /
// leaves from shared memory as a dense [box_rows][box_cols] array.  Out-of-range parts of a tile are not
// written (stores) / zero-filled (loads), so ragged last tiles need no special path.  The driver entry point
// is looked up through the runtime (no link-time dependency on libcuda).
inline bool make_tensor_map_2d_f32(CUtensorMap *map, const void *base, uint64_t rows, uint64_t cols,
                                   uint32_t box_rows, uint32_t box_cols) {
  typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                               const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                               CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess || !fn)
      return false;
    encode = reinterpret_cast<EncodeFn>(fn);
  }
  if ((cols * 4) % 16 != 0 || (reinterpret_cast<uintptr_t>(base) & 15) != 0 || box_rows > 256 || box_cols > 256) return false;
  const cuuint64_t gdim[2] = {cols, rows};
  const cuuint64_t gstride[1] = {cols * 4};
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base), gdim, gstride, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// One thread: store the dense shared-memory tile at tensor coordinates (col0, row0) (UTMASTG); completion is
// tracked by the thread's bulk async-group.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, const void *smem_src, int col0, int row0) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
               "r"(smem_u32(smem_src)), "r"(col0), "r"(row0)
               : "memory");
}
// One thread: load the tile at (col0, row0) into shared memory (UTMALDG); completes on the mbarrier.
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, int col0, int row0, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(col0), "r"(row0)
      : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---------------------------------------------------------------- fences
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---------------------------------------------------------------- TMEM
// One full warp allocates `ncols` (power of two >= 32) columns; the base address
// lands in *slot (shared memory).
__device__ __forceinline__ void tmem_alloc(uint32_t *slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem] . B[smem]; issued by ONE thread for the whole CTA.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive once all tcgen05 ops previously issued by this thread are done.
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// 32 lanes x 32 bit: thread i of warp w reads TMEM lane 32*(w%4)+i, 16 consecutive columns.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

__host__ __device__ constexpr uint32_t make_idesc_bf16(int n, bool a_mn_major, bool b_mn_major, int m = 128) {
  return make_idesc_16(n, a_mn_major, b_mn_major, 1u, m);
}

// ---------------------------------------------------------------- packing
// F16 = true: IEEE half operands (11-bit significand), false: bfloat16 (8-bit).
// The half conversions SATURATE (cvt.rn.satfinite, one F2FP like the plain conversion): an activation
// beyond +-65504 -- GELU(hidden), att @ v, post-ReLU RSB maps are unbounded with trained weights -- becomes
// +-65504 instead of an infinity that would turn the whole clip into NaNs in the next UMMA.  NaN stays NaN.
template <bool F16>
__device__ __forceinline__ uint32_t pack16x2(float lo, float hi) {
  if constexpr (F16) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
  } else {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&h);
  }
}
template <bool F16>
__device__ __forceinline__ uint4 pack16x8(const float *v) {
  return make_uint4(pack16x2<F16>(v[0], v[1]), pack16x2<F16>(v[2], v[3]), pack16x2<F16>(v[4], v[5]),
                    pack16x2<F16>(v[6], v[7]));
}
template <bool F16>
__device__ __forceinline__ unsigned short to16(float v) {
  if constexpr (F16) {
    unsigned short h;
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(v));
    return h;
  } else {
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    return *reinterpret_cast<unsigned short *>(&h);
  }
}

// byte offset of element (r, k) in a core-matrix interleaved tile
__host__ __device__ constexpr uint32_t cm_offset(uint32_t r, uint32_t k, uint32_t rs, uint32_t cs) {
  return (r >> 3) * rs + (k >> 3) * cs + (r & 7) * 16 + (k & 7) * 2;
}

// 16-byte asynchronous global -> shared copy (LDGSTS)
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

}  // namespace tc
}  // namespace otp
