// tcgen05 / TMEM / mbarrier building blocks (sm_100a inline PTX) shared by the tensor-core kernels.
//
// Operand layouts are the canonical K-major, no-swizzle ("interleave") UMMA layout: 8x16-byte core
// matrices; inside a core matrix the 8 rows sit at consecutive 16-byte lines; LBO is the byte distance
// between core matrices that are adjacent in K, SBO the byte distance between 8-row groups along M/N:
//     byte_offset(row, k) = (row / 8) * SBO + (row % 8) * 16 + (k / 8) * LBO + (k % 8) * 2      (16-bit elements)
// One tcgen05.mma of kind::f16 consumes K = 16 elements = two core matrices LBO apart.
#pragma once
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cstdint>

namespace orcai {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// 64-bit shared-memory matrix descriptor (SM100 format: version field = 1, no swizzle, base offset 0)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);            // start address, bits [0,14)
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;   // leading byte offset, bits [16,30)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;   // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                              // descriptor version (Blackwell)
  return d;                                            // layout_type bits [61,64) = 0: SWIZZLE_NONE
}

// 32-bit instruction descriptor for kind::f16, fp32 accumulate, both operands K-major
// fmt: 0 = fp16, 1 = bf16
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N, int fmt) {
  return (1u << 4)                      // c_format = F32
         | ((uint32_t)fmt << 7)         // a_format
         | ((uint32_t)fmt << 10)        // b_format
         | ((uint32_t)(N >> 3) << 17)   // n_dim
         | ((uint32_t)(M >> 4) << 24);  // m_dim
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// make generic-proxy shared-memory writes visible to the async proxy (the tensor core reads operands through it)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Bring-up aid (compile with -DORCAI_TRAP_INFO, run with ORCAI_B200_TRAPINFO=1; net_tc.cu): a wait that gives up records
// {magic, blockIdx.x, threadIdx.x, barrier shared address, parity} in mapped host memory before it traps, and
// orcai_last_error appends the record.  Compiled out by default: the extra live values of the cold path cost registers
// (spills in conv0_mma_kernel and lstm_rec_tc_kernel, +0.5 ms per hour of audio when it was always on).
#ifdef ORCAI_TRAP_INFO
static __device__ unsigned int* g_trap_info = nullptr;   // one copy per translation unit (no relocatable device code)
#endif

// Wait for the phase with the given parity.  Bounded: a broken pipeline traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (!done && spin > (1u << 26)) {
#ifdef ORCAI_TRAP_INFO
      unsigned int* ti = g_trap_info;
      if (ti != nullptr && atomicCAS(ti, 0u, 0x7241u) == 0u) {
        ti[1] = blockIdx.x; ti[2] = threadIdx.x; ti[3] = addr; ti[4] = parity;
        __threadfence_system();
      }
#endif
      __trap();
    }
  }
}

// TMEM allocation: one full warp executes; the base address lands in *slot (shared memory)
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot) {
  static_assert(NCOLS >= 32 && NCOLS <= 512 && (NCOLS & (NCOLS - 1)) == 0, "TMEM columns: power of two in [32, 512]");
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(NCOLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread
__device__ __forceinline__ void mma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// TMEM -> registers: this warp's 32 lanes (lane = accumulator row), 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// 16-bit storage types
template <typename H> struct half_traits;
template <> struct half_traits<__half> {
  static constexpr int fmt = 0;
  static __device__ __forceinline__ __half from_float(float x) { return __float2half_rn(x); }
  static __device__ __forceinline__ float to_float(__half h) { return __half2float(h); }
};
template <> struct half_traits<__nv_bfloat16> {
  static constexpr int fmt = 1;
  static __device__ __forceinline__ __nv_bfloat16 from_float(float x) { return __float2bfloat16_rn(x); }
  static __device__ __forceinline__ float to_float(__nv_bfloat16 h) { return __bfloat162float(h); }
};

// 256-bit global store (sm_100: STG.E.256), p 32-byte aligned.  A thread that owns a row (pixel / GEMM row) whose neighbours' rows
// are far away fills a whole 32-byte sector per instruction with it; with 16-byte stores every sector is written half at a time.
__device__ __forceinline__ void st_global_v8(float* p, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]),
               "f"(v[6]), "f"(v[7])
               : "memory");
}

}  // namespace tc
}  // namespace orcai
