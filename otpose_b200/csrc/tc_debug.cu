// Self-test of the tcgen05 plumbing: one CTA computes D[128, N] = A . B^T from
// caller-built shared-memory operand images with caller-chosen descriptor fields.
// Used by tests/test_gpu_tc.py to pin the descriptor semantics (K-major and
// MN-major views of the core-matrix interleaved layout) against a CPU GEMM.
#include "common.cuh"
#include "tc_common.cuh"

namespace otp {

__global__ void __launch_bounds__(128)
umma_selftest_kernel(const uint4 *__restrict__ a_img, int a_bytes, const uint4 *__restrict__ b_img, int b_bytes,
                     float *__restrict__ d, int n, int ksteps, uint32_t a_off, uint32_t a_lbo, uint32_t a_sbo,
                     uint32_t a_kstep, uint32_t b_off, uint32_t b_lbo, uint32_t b_sbo, uint32_t b_kstep,
                     int a_mn, int b_mn, int repeat) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  uint8_t *as = smem;
  uint8_t *bs = smem + ((a_bytes + 1023) / 1024) * 1024;
  for (int i = threadIdx.x; i < a_bytes / 16; i += 128) reinterpret_cast<uint4 *>(as)[i] = a_img[i];
  for (int i = threadIdx.x; i < b_bytes / 16; i += 128) reinterpret_cast<uint4 *>(bs)[i] = b_img[i];
  if (threadIdx.x == 0) {
    tc::mbar_init(&bar, 1);
    tc::fence_mbar_init();
  }
  if (threadIdx.x < 32) tc::tmem_alloc(&tmem_slot, 256);
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = tc::make_idesc_bf16(n, a_mn != 0, b_mn != 0);
  uint32_t parity = 0;
  for (int rep = 0; rep < repeat; ++rep) {
    if (threadIdx.x == 0) {
      for (int s = 0; s < ksteps; ++s) {
        uint64_t ad = tc::make_desc(tc::smem_u32(as) + a_off + s * a_kstep, a_lbo, a_sbo);
        uint64_t bd = tc::make_desc(tc::smem_u32(bs) + b_off + s * b_kstep, b_lbo, b_sbo);
        tc::umma_bf16(tmem, ad, bd, idesc, (s > 0 || rep > 0) ? 1u : 0u);
      }
      tc::umma_commit(&bar);
    }
    tc::mbar_wait(&bar, parity);
    parity ^= 1;
  }
  tc::tc_fence_after();
  const int warp = threadIdx.x >> 5, row = threadIdx.x;
  for (int c0 = 0; c0 < n; c0 += 8) {
    float v[8];
    tc::tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) d[(size_t)row * n + c0 + i] = v[i];
  }
  tc::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tc::tmem_dealloc(tmem, 256);
}


// UMMA issue/execute rate probe: one warp per CTA issues `reps` x `ksteps` M128 x N x K16 UMMAs
// (operands: zeroed shared memory, K-major) and measures clock64 from first issue to completion.
__global__ void __launch_bounds__(128)
umma_rate_kernel(int n, int ksteps, int reps, long long *__restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t rs = (uint32_t)ksteps * 2 * 128;            // row-group stride of a K = 16*ksteps tile
  const uint32_t a_bytes = 16 * rs, b_bytes = (uint32_t)(n / 8) * rs;
  for (uint32_t i = threadIdx.x; i < (a_bytes + b_bytes) / 16; i += 128)
    reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    tc::mbar_init(&bar, 1);
    tc::fence_mbar_init();
  }
  if (threadIdx.x < 32) tc::tmem_alloc(&tmem_slot, 256);
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  if (threadIdx.x < 32) {
    const uint32_t tmem = tmem_slot, aa = tc::smem_u32(smem), bb = aa + a_bytes;
    const uint32_t idesc = tc::make_idesc_bf16(n, false, false);
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r)
      for (int s = 0; s < ksteps; ++s) {
        const uint64_t ad = tc::make_desc(aa + s * 256, 128, rs), bd = tc::make_desc(bb + s * 256, 128, rs);
        asm volatile(
            "{\n\t.reg .pred p, q;\n\t"
            "elect.sync _|q, 0xffffffff;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"((uint32_t)(s > 0))
            : "memory");
      }
    const long long t1 = clock64();
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
        ::"r"(tc::smem_u32(&bar))
        : "memory");
    tc::mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if (threadIdx.x == 0) {
      cycles[2 * blockIdx.x] = t1 - t0;       // issue
      cycles[2 * blockIdx.x + 1] = t2 - t0;   // issue + drain
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tc::tmem_dealloc(tmem_slot, 256);
}

}  // namespace otp

extern "C" int otp_debug_umma_gemm(const void *a_img, int a_bytes, const void *b_img, int b_bytes, float *d,
                                   int n, int ksteps, unsigned a_off, unsigned a_lbo, unsigned a_sbo,
                                   unsigned a_kstep, unsigned b_off, unsigned b_lbo, unsigned b_sbo,
                                   unsigned b_kstep, int a_mn_major, int b_mn_major, int repeat,
                                   otp_stream_t stream) {
  OTP_REQUIRE(a_img && b_img && d && n >= 16 && n <= 256 && n % 16 == 0 && ksteps > 0 && repeat > 0);
  OTP_REQUIRE(a_bytes % 16 == 0 && b_bytes % 16 == 0);
  size_t smem = (size_t)((a_bytes + 1023) / 1024) * 1024 + b_bytes;
  OTP_REQUIRE(smem <= 200 * 1024);
  cudaFuncSetAttribute(otp::umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  otp::umma_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(
      static_cast<const uint4 *>(a_img), a_bytes, static_cast<const uint4 *>(b_img), b_bytes, d, n, ksteps, a_off,
      a_lbo, a_sbo, a_kstep, b_off, b_lbo, b_sbo, b_kstep, a_mn_major, b_mn_major, repeat);
  return otp::check_launch("umma_selftest_kernel");
}

extern "C" int otp_debug_umma_rate(int n, int ksteps, int reps, int ctas, long long *cycles_host) {
  using namespace otp;
  OTP_REQUIRE(n >= 16 && n <= 256 && n % 16 == 0 && ksteps >= 1 && ksteps <= 9 && reps >= 1 && ctas >= 1 &&
              ctas <= 1024 && cycles_host != nullptr);
  const size_t smem = (size_t)(16 + n / 8) * ksteps * 256;
  cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  long long *dev = nullptr;
  if (cudaMalloc(&dev, sizeof(long long) * 2 * ctas) != cudaSuccess) return OTP_ERR_CUDA;
  umma_rate_kernel<<<ctas, 128, smem>>>(n, ksteps, reps, dev);
  cudaError_t e = cudaMemcpy(cycles_host, dev, sizeof(long long) * 2 * ctas, cudaMemcpyDeviceToHost);
  cudaFree(dev);
  if (e != cudaSuccess) {
    set_error("otp_debug_umma_rate: %s", cudaGetErrorString(e));
    return OTP_ERR_CUDA;
  }
  return OTP_OK;
}
