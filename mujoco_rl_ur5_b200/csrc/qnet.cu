// libgrasp_qnet.so — forward pass of the pixel-wise grasp Q-network (reference: Modules.py:308-311 MULTIDISCRETE_RESNET =
// Perception_Module :159-193 + Grasping_Module_multidiscrete :243-287, BasicBlock :92-142) over a batch of environments.
//
// bf16 activations (NHWC) and weights, fp32 accumulation.  The 3x3 / 1x1 convolutions with Cin >= 64 (>99.9 % of the MACs) run
// as implicit GEMMs on the 5th-generation tensor cores: tcgen05.mma (kind::f16, M=128, N=64/128, K=16) with the accumulator in
// TMEM, operands gathered into shared memory in the canonical K-major no-swizzle core-matrix layout, completion tracked with
// tcgen05.commit -> mbarrier, epilogue through tcgen05.ld.  BatchNorm uses per-image batch statistics because the reference
// never puts the policy net in eval() and forwards one image at a time.
//
// C-ABI: see include/grasp_qnet.h.  No CPU fallback.
#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched with cudaGetDriverEntryPoint, libcuda is not linked)
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/grasp_qnet.h"
#include "qnet_plan.h"

typedef __nv_bfloat16 bf16;

static thread_local char q_err[256] = "";
extern "C" const char* gq_last_error(void) { return q_err; }
#define QCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { snprintf(q_err, sizeof q_err, "CUDA error: %s", cudaGetErrorString(e_)); return -2; } } while (0)

// ------------------------------------------------------------------------------------------------ small PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0, addr = smem_u32(bar);
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T ; single-thread issue
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t of the warp receives row (lane base + t)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}
// shared-memory matrix descriptor, K-major, no swizzle: core matrix = 8 rows x 16 B stored contiguously (128 B);
// LBO = byte distance between the two 16-byte K-chunks of one MMA, SBO = byte distance between 8-row groups (kept for reference;
// the convolution uses the 128-byte-swizzled layout below)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // base offset 0, lbo mode 0, layout type 0 = SWIZZLE_NONE
}
// shared-memory matrix descriptor, K-major, SWIZZLE_128B: one tile row = 128 contiguous bytes (64 bf16 along K) whose eight 16-byte
// chunks are XOR-ed with (row % 8); 8 rows form a 1024-byte atom (SBO), the tile base is 1024-byte aligned.  A K-step of 16
// elements advances the start address by 32 bytes inside the atom.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;            // leading byte offset: unused for swizzled K-major
  d |= (uint64_t)(1024 >> 4) << 32;  // stride byte offset: 8 rows x 128 B
  d |= (uint64_t)1 << 46;            // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;            // layout type SWIZZLE_128B
  return d;
}
// instruction descriptor kind::f16: D = F32, A = B = BF16, both K-major, shape M x N (K = 16)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------ implicit-GEMM convolution on tcgen05
// x  [B,H,W,Cin] bf16 (NHWC), w [Cout][ks*ks][Cin] bf16, y [B,H,W,Cout] f32 (+ bias), stats [B,Cout,2] f32 += (sum, sum of squares)
// grid (ceil(H*W/128), Cout/BLOCK_N, B), 128 threads.  One CTA = 128 output pixels x BLOCK_N output channels of one image.
#define BM 128
#define BK 64
// 16-byte asynchronous global -> shared copy (LDGSTS); src_bytes = 0 writes zeros (convolution padding, rows past the image)
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async16_s(uint32_t dst_smem_addr, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem_addr), "l"(src), "r"(src_bytes) : "memory");
}
// TMA: one thread copies a whole [rows x 64 bf16] box of a 2-D tensor into a 128-byte-swizzled shared-memory tile; completion is
// reported to the mbarrier as transferred bytes (complete_tx), so the same "full" barrier collects the cp.async arrivals and the TMA
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem_addr, const CUtensorMap* tmap, int x, int y, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst_smem_addr), "l"((uint64_t)tmap), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}
// .cg: cache in L2 only (weight tiles are read once per CTA; keeping them out of L1 leaves it to the activation rows the 9 taps re-read)
__device__ __forceinline__ void cp_async16_cg(uint32_t dst_smem_addr, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem_addr), "l"(src) : "memory");
}
// Column sums of a 32 x 32 tile held one row per lane (v[c] = element (lane, c)): on return lane j holds sum_lanes v[j].
// Recursive halving: at distance s the lanes with bit s set keep the upper s columns of what they still hold and hand the lower s to
// their partner (and vice versa), so 16 + 8 + 4 + 2 + 1 = 31 shuffles replace 32 x 5; the summation tree is fixed (deterministic).
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int k = 0; k < s; k++) {
      const float send = up ? v[k] : v[k + s];
      const float keep = up ? v[k + s] : v[k];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// arrive on the mbarrier once all cp.async issued so far by this thread have landed (does not bump the expected count)
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- non-persistent convolution kernel (GQ_PERSIST=0; the default until r02a, kept as the A/B partner of the persistent TMA kernel
// below and as the path for drivers without cuTensorMapEncodeIm2col).  Same tiling, shared-memory layout, barrier protocol and MMA
// issue as the first version of round 1 (removed in r02); what changed is what the ncu capture of v1 showed to be the limit (profiles/r01h_conv_tc_ncu_summary.txt: tensor pipe 17-23 % active,
// issue slots 30-44 % busy with 1.25 resident warps per scheduler = the four producer warps executing ~700 instructions per k-step):
//  * producer address arithmetic is hoisted out of the k-loop: every thread's 8 A rows and BLOCK_N/16 B rows are an arithmetic
//    progression in global memory (row stride 16 pixels / 16 output channels) and in shared memory (2048 B: 16 rows x 128 B, the
//    128-byte swizzle term depends on row % 8 only and is therefore the same for all rows of a thread); the zero-padding test of
//    each (row, tap) is a bit of a per-row mask computed once; tap / channel-chunk / stage / phase are running counters (no divisions)
//  * the per-channel BatchNorm partial sums of a 32 x 32 accumulator block use a transposing reduction (31 shuffles instead of 160)
//    and are stored as one coalesced 256-byte row per warp
//  * EPI == 1 (BasicBlock tail): the 1x1 shortcut convolution adds the normalised main branch and applies ReLU in its epilogue,
//    out = relu(resid * scale + shift + conv(x) + bias) as bf16, which removes the fp32 round trip of the shortcut and one
//    BN-apply kernel per block (BasicBlock.forward, Modules.py:128-142)
//  * NPW = 4 or 8 producer warps (GQ_NPW): 8 halves the copies per thread and lets warps 4-7 take half of the epilogue columns (a warp
//    may read the TMEM lane quarter warp % 4, so warps w and w + 4 share rows and split columns)
//  * TMAB = 1 (default; GQ_TMA=0 disables): the weight tile of a k-step (two thirds of the bytes at BLOCK_N = 256) is one TMA box issued by one thread
//    instead of BLOCK_N / RS cp.async per producer thread; the activation gather stays on cp.async (its rows are not a box)
template <int BLOCK_N, int STAGES, int EPI, int NPW, int TMAB>
__global__ void __launch_bounds__(32 * (NPW + 1)) k_conv_tc(const __grid_constant__ CUtensorMap tmap_w, const bf16* __restrict__ x,
                                                            const bf16* __restrict__ w, const float* __restrict__ bias,
                                                 float* __restrict__ y, float* __restrict__ stats, const float* __restrict__ resid,
                                                 const float2* __restrict__ scale_shift, bf16* __restrict__ out, int H, int W, int Cin, int Cout,
                                                 int ks, int cg_weights) {
  constexpr int A_STAGE = BM * BK * 2, B_STAGE = BLOCK_N * BK * 2;  // bytes
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_STAGE;
  uint64_t* full = (uint64_t*)(smem + STAGES * (A_STAGE + B_STAGE));  // [STAGES] copies landed (32 NPW producer arrivals)
  uint64_t* empty = full + STAGES;                                    // [STAGES] MMAs that read the stage are done (1 commit)
  uint64_t* accbar = empty + STAGES;                                  // accumulator complete
  uint32_t* tmem_slot = (uint32_t*)(accbar + 1);
  constexpr int RS = 4 * NPW;  // row step of the copy mapping (qnet_plan.h): producer threads / 8
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int HW = H * W, m0 = blockIdx.x * BM, n0 = blockIdx.y * BLOCK_N, b = blockIdx.z;
  const int pad = ks / 2, taps = ks * ks, kchunks = Cin / BK, nk = taps * kchunks;
  if (tid == 0) {
    for (int i = 0; i < STAGES; i++) { mbar_init(&full[i], 32 * NPW + TMAB); mbar_init(&empty[i], 1); }
    mbar_init(accbar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, BLOCK_N);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tacc = *tmem_slot;
  constexpr uint32_t idesc = make_idesc(BM, BLOCK_N);
  if (warp < NPW) {
    // ---- producers.  Thread -> 16-byte chunk tid % 8 of rows tid / 8 + RS i (consecutive lanes take consecutive chunks of one
    // row: a warp-level cp.async reads 4 full 128-byte lines and writes 4 swizzled, conflict-free shared-memory rows).  The plan
    // (qnet_plan.h) holds the per-thread constants and the running tap / chunk / stage counters.
    ConvPlan p;
    conv_plan_init<RS>(p, tid, m0, n0, H, W, Cin, ks);
    const uint32_t sA_u = smem_u32(sA) + p.dstoff, sB_u = smem_u32(sB) + p.dstoff;
    const char* xb = (const char*)(x + (size_t)b * HW * Cin);
    const char* wb = (const char*)w;
    for (int kn = 0; kn < nk; kn++) {
      if (p.round > 0) mbar_wait(&empty[p.sn], (uint32_t)(p.round - 1) & 1u);  // the MMAs of k-step kn - STAGES have read the stage
      const uint32_t tbit = 1u << p.tap;
      const char* ap = xb + conv_plan_a(p, W, Cin);
      const uint32_t da = sA_u + (uint32_t)(p.sn * A_STAGE);
#pragma unroll
      for (int i = 0; i < BM / RS; i++) {
        const bool ok = (p.vmask[i] & tbit) != 0;
        cp_async16_s(da + i * (RS * 128), ok ? (const void*)ap : (const void*)xb, ok ? 16u : 0u);
        ap += p.a_stride;
      }
      if constexpr (TMAB) {
        if (tid == 0) {  // box = rows n0 .. n0 + BLOCK_N - 1, columns 64 kn .. 64 kn + 63 of w viewed as [Cout][taps * Cin]
          mbar_arrive_expect_tx(&full[p.sn], (uint32_t)B_STAGE);
          tma_load_2d(smem_u32(sB) + (uint32_t)(p.sn * B_STAGE), &tmap_w, kn * BK, n0, &full[p.sn]);
        }
      } else {
        const char* bp = wb + conv_plan_b(p, kn);
        const uint32_t db = sB_u + (uint32_t)(p.sn * B_STAGE);
#pragma unroll
        for (int i = 0; i < BLOCK_N / RS; i++) {
          if (cg_weights) cp_async16_cg(db + i * (RS * 128), bp); else cp_async16_s(db + i * (RS * 128), bp, 16u);
          bp += p.b_stride;
        }
      }
      cp_async_mbar_arrive(&full[p.sn]);
      conv_plan_next(p, STAGES, kchunks, pad);
    }
  } else if (lane == 0) {  // warp NPW: the MMA issuer
    int s = 0; uint32_t ph = 0;
    for (int kb = 0; kb < nk; kb++) {
      mbar_wait(&full[s], ph);
      fence_proxy_async();  // generic-proxy (cp.async) smem writes -> visible to the tensor core (async proxy)
      tc_fence_after();
      const uint32_t a_base = smem_u32(sA + s * A_STAGE), b_base = smem_u32(sB + s * B_STAGE);
#pragma unroll
      for (int k = 0; k < BK / 16; k++) {  // UMMA_K = 16 bf16 = 32 bytes inside the swizzle atom
        uint64_t adesc = make_smem_desc_sw128(a_base + k * 32);
        uint64_t bdesc = make_smem_desc_sw128(b_base + k * 32);
        umma_bf16(tacc, adesc, bdesc, idesc, (kb > 0 || k > 0) ? 1u : 0u);
      }
      umma_commit(&empty[s]);
      if (kb == nk - 1) umma_commit(accbar);
      if (++s == STAGES) { s = 0; ph ^= 1u; }
    }
  }
  if (warp < NPW) {
    mbar_wait(accbar, 0);
    tc_fence_after();
    // ---- epilogue: lane t of a warp with warp % 4 == q holds accumulator row 32 q + t; with 8 warps, warps 4-7 take the upper half of
    // the columns
    const int quarter = warp & 3;
    const int m = m0 + quarter * 32 + lane;
    const bool mvalid = m < HW;
    const size_t rowoff = ((size_t)b * HW + m) * Cout + n0;
    constexpr int CB_PER_WARP = BLOCK_N / (NPW / 4);
    const int cb0 = (warp >> 2) * CB_PER_WARP;
    for (int cb = cb0; cb < cb0 + CB_PER_WARP; cb += 32) {
      float v[32];
      tmem_ld32(tacc + ((uint32_t)(quarter * 32) << 16) + cb, v);
      if (bias) {
        const float4* b4 = (const float4*)(bias + n0 + cb);
#pragma unroll
        for (int i = 0; i < 8; i++) { const float4 bv = __ldg(b4 + i); v[4 * i] += bv.x; v[4 * i + 1] += bv.y; v[4 * i + 2] += bv.z; v[4 * i + 3] += bv.w; }
      }
      if constexpr (EPI == 0) {
        if (mvalid) {
          float4* dst = (float4*)(y + rowoff + cb);
#pragma unroll
          for (int i = 0; i < 8; i++) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
        if (stats) {
          float q[32];
#pragma unroll
          for (int i = 0; i < 32; i++) { v[i] = mvalid ? v[i] : 0.f; q[i] = v[i] * v[i]; }
          const float s1 = warp_colsum32(v, lane), s2 = warp_colsum32(q, lane);
          // deterministic: every (CTA, warp) writes its own partial row, k_bn_reduce adds them in a fixed order
          float2* pp = (float2*)(stats + ((((size_t)b * gridDim.x + blockIdx.x) * 4 + quarter) * Cout + n0 + cb) * 2);
          pp[lane] = make_float2(s1, s2);
        }
      } else {
        if (mvalid) {
          const float4* r4 = (const float4*)(resid + rowoff + cb);
          const float4* ss4 = (const float4*)(scale_shift + (size_t)b * Cout + n0 + cb);  // (scale, shift) pairs
          __nv_bfloat162 o[16];
#pragma unroll
          for (int i = 0; i < 8; i++) {
            const float4 r = __ldg(r4 + i);
            const float4 sa = __ldg(ss4 + 2 * i), sb = __ldg(ss4 + 2 * i + 1);  // columns 4i, 4i+1 | 4i+2, 4i+3
            const float o0 = fmaxf(r.x * sa.x + sa.y + v[4 * i], 0.f), o1 = fmaxf(r.y * sa.z + sa.w + v[4 * i + 1], 0.f);
            const float o2 = fmaxf(r.z * sb.x + sb.y + v[4 * i + 2], 0.f), o3 = fmaxf(r.w * sb.z + sb.w + v[4 * i + 3], 0.f);
            o[2 * i] = __floats2bfloat162_rn(o0, o1); o[2 * i + 1] = __floats2bfloat162_rn(o2, o3);
          }
          uint4* dst = (uint4*)(out + rowoff + cb);
#pragma unroll
          for (int i = 0; i < 4; i++) dst[i] = ((const uint4*)o)[i];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tacc, BLOCK_N);
}
// epilogue of one 128 x BLOCK_N accumulator (TMEM columns tcol .. tcol + BLOCK_N - 1) for the warp owning lane quarter `quarter`:
// the same arithmetic as the epilogue of k_conv_tc
template <int BLOCK_N, int EPI>
__device__ __forceinline__ void conv_epilogue_tile(uint32_t tcol, int quarter, int lane, int mt, int n0, int b, int gx, int HW, int Cout,
                                                   const float* __restrict__ bias, float* __restrict__ y, float* __restrict__ stats,
                                                   const float* __restrict__ resid, const float2* __restrict__ scale_shift, bf16* __restrict__ out) {
  const int m = mt * BM + quarter * 32 + lane;
  const bool mvalid = m < HW;
  const size_t rowoff = ((size_t)b * HW + m) * Cout + n0;
  for (int cb = 0; cb < BLOCK_N; cb += 32) {
    float v[32];
    tmem_ld32(tcol + ((uint32_t)(quarter * 32) << 16) + (uint32_t)cb, v);
    if (bias) {
      const float4* b4 = (const float4*)(bias + n0 + cb);
#pragma unroll
      for (int j = 0; j < 8; j++) { const float4 bv = __ldg(b4 + j); v[4 * j] += bv.x; v[4 * j + 1] += bv.y; v[4 * j + 2] += bv.z; v[4 * j + 3] += bv.w; }
    }
    if constexpr (EPI == 0) {
      if (mvalid) {
        float4* dst = (float4*)(y + rowoff + cb);
#pragma unroll
        for (int j = 0; j < 8; j++) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
      if (stats) {
        float q[32];
#pragma unroll
        for (int j = 0; j < 32; j++) { v[j] = mvalid ? v[j] : 0.f; q[j] = v[j] * v[j]; }
        const float s1 = warp_colsum32(v, lane), s2 = warp_colsum32(q, lane);
        float2* pp = (float2*)(stats + ((((size_t)b * gx + mt) * 4 + quarter) * Cout + n0 + cb) * 2);
        pp[lane] = make_float2(s1, s2);
      }
    } else {
      if (mvalid) {
        const float4* r4 = (const float4*)(resid + rowoff + cb);
        const float4* ss4 = (const float4*)(scale_shift + (size_t)b * Cout + n0 + cb);
        __nv_bfloat162 o[16];
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const float4 r = __ldg(r4 + j);
          const float4 sa = __ldg(ss4 + 2 * j), sb = __ldg(ss4 + 2 * j + 1);
          const float o0 = fmaxf(r.x * sa.x + sa.y + v[4 * j], 0.f), o1 = fmaxf(r.y * sa.z + sa.w + v[4 * j + 1], 0.f);
          const float o2 = fmaxf(r.z * sb.x + sb.y + v[4 * j + 2], 0.f), o3 = fmaxf(r.w * sb.z + sb.w + v[4 * j + 3], 0.f);
          o[2 * j] = __floats2bfloat162_rn(o0, o1); o[2 * j + 1] = __floats2bfloat162_rn(o2, o3);
        }
        uint4* dst = (uint4*)(out + rowoff + cb);
#pragma unroll
        for (int j = 0; j < 4; j++) dst[j] = ((const uint4*)o)[j];
      }
    }
  }
}

// ---- DEFAULT since r02a: the persistent kernel fed by TMA only (whole forward 477 vs 439 TFLOP/s, profiles/r02a_qnet_sweep.txt; a
// second persistent variant with a cp.async activation gather measured 409 and was deleted).  One CTA per SM walks a static
// round-robin list of (m tile, n tile, image) tiles; the accumulator is double-buffered in TMEM (2 x BLOCK_N columns) and the epilogue
// has its own 4 warps, so the TMEM read-out, the stores and the BatchNorm partials of tile i overlap the main loop of tile i + 1.  The
// activation operand is an **im2col-mode** tensor map over x [N][H][W][C] (cuTensorMapEncodeIm2col: pixel box corners (-pad, -pad) /
// (pad - (ks - 1), ...), 64 channels per pixel, 128 pixels per column, 128-byte swizzle): one
// `cp.async.bulk.tensor.4d...im2col` per k-step loads the 128 consecutive output positions of the tile for filter tap (offset_w, offset_h)
// with hardware zero fill at the image border - the same [128 pixels x 64 channels] swizzled tile the cp.async gather builds (semantics
// as used by CUTLASS' sm100 implicit-GEMM collective: start coordinate = first output pixel + lower corner, tap passed as offsets).
// One thread issues both TMAs of a k-step after a single arrive.expect_tx(A + B bytes) on the stage's full barrier (count 1).
// Roles: warp 0 lane 0 producer, warp 1 lane 0 MMA issuer, warps 2-5 epilogue (lane quarter warp % 4).
__device__ __forceinline__ void tma_load_im2col_4d(uint32_t dst_smem_addr, const CUtensorMap* tmap, int c, int w, int h, int n, uint16_t off_w,
                                                   uint16_t off_h, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
               ::"r"(dst_smem_addr), "l"((uint64_t)tmap), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h) : "memory");
}
template <int BLOCK_N, int STAGES, int EPI>
__global__ void __launch_bounds__(192) k_conv_tc_tma(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x,
                                                     const float* __restrict__ bias, float* __restrict__ y, float* __restrict__ stats,
                                                     const float* __restrict__ resid, const float2* __restrict__ scale_shift, bf16* __restrict__ out,
                                                     int H, int W, int Cin, int Cout, int ks, int gx, int gy, int ntiles) {
  constexpr int A_STAGE = BM * BK * 2, B_STAGE = BLOCK_N * BK * 2;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_STAGE;
  uint64_t* full = (uint64_t*)(smem + STAGES * (A_STAGE + B_STAGE));
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = (uint32_t*)(acc_empty + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int HW = H * W, pad = ks / 2, taps = ks * ks, kchunks = Cin / BK, nk = taps * kchunks;
  if (tid == 0) {
    for (int i = 0; i < STAGES; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; i++) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 128); }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 2 * BLOCK_N);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tacc = *tmem_slot;
  constexpr uint32_t idesc = make_idesc(BM, BLOCK_N);
  if (warp == 0) {
    if (lane == 0) {
      int sn = 0, round = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int mt = t % gx, nt = (t / gx) % gy, b = t / (gx * gy);
        const int m0 = mt * BM, n0 = nt * BLOCK_N;
        const int oh0 = m0 / W, ow0 = m0 - oh0 * W;
        int tap = 0, kc = 0;
        for (int kn = 0; kn < nk; kn++) {
          if (round > 0) mbar_wait(&empty[sn], (uint32_t)(round - 1) & 1u);
          mbar_arrive_expect_tx(&full[sn], (uint32_t)(A_STAGE + B_STAGE));
          tma_load_im2col_4d(smem_u32(sA) + (uint32_t)(sn * A_STAGE), &tmap_x, kc * BK, ow0 - pad, oh0 - pad, b, (uint16_t)(tap % ks), (uint16_t)(tap / ks),
                             &full[sn]);
          tma_load_2d(smem_u32(sB) + (uint32_t)(sn * B_STAGE), &tmap_w, kn * BK, n0, &full[sn]);
          if (++sn == STAGES) { sn = 0; round++; }
          if (++kc == kchunks) { kc = 0; tap++; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int s = 0; uint32_t ph = 0; int i = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x, i++) {
        const int a = i & 1;
        mbar_wait(&acc_empty[a], (uint32_t)((i >> 1) & 1) ^ 1u);
        tc_fence_after();
        const uint32_t dacc = tacc + (uint32_t)(a * BLOCK_N);
        for (int kb = 0; kb < nk; kb++) {
          mbar_wait(&full[s], ph);  // both operands arrive through the async proxy (TMA): no generic-proxy fence needed
          tc_fence_after();
          const uint32_t a_base = smem_u32(sA + s * A_STAGE), b_base = smem_u32(sB + s * B_STAGE);
#pragma unroll
          for (int k = 0; k < BK / 16; k++)
            umma_bf16(dacc, make_smem_desc_sw128(a_base + k * 32), make_smem_desc_sw128(b_base + k * 32), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty[s]);
          if (kb == nk - 1) umma_commit(&acc_full[a]);
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else {
    const int quarter = warp & 3;
    int i = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, i++) {
      const int mt = t % gx, nt = (t / gx) % gy, b = t / (gx * gy);
      const int a = i & 1;
      mbar_wait(&acc_full[a], (uint32_t)(i >> 1) & 1u);
      tc_fence_after();
      conv_epilogue_tile<BLOCK_N, EPI>(tacc + (uint32_t)(a * BLOCK_N), quarter, lane, mt, nt * BLOCK_N, b, gx, HW, Cout, bias, y, stats, resid, scale_shift, out);
      tc_fence_before();
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acc_empty[a])) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tacc, 2 * BLOCK_N);
}

// per-image BatchNorm (training mode, own statistics) folded into one multiply-add per channel: (scale, shift) from (sum, sum of squares)
__global__ void k_bn_scale_shift(const float* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta, float2* __restrict__ ss,
                                 int B, int C, int HW, float eps) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  int c = i % C;
  float s1 = stats[(size_t)i * 2], s2 = stats[(size_t)i * 2 + 1];
  float mean = s1 / HW, var = fmaxf(s2 / HW - mean * mean, 0.f);  // same expressions as k_bn_act
  float scale = rsqrtf(var + eps) * gamma[c];
  ss[i] = make_float2(scale, beta[c] - mean * scale);
}

// per-image batch-norm statistics from the per-(tile, warp) partials of k_conv_tc, summed in a fixed order (double accumulation)
__global__ void k_bn_reduce(const float* __restrict__ part, float* __restrict__ stats, int B, int nparts, int C) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  int b = i / C, c = i % C;
  // one accumulator pair in part order (the order is part of the result); the unrolled body keeps 8 independent 8-byte loads in flight
  // (r01j: a 4-accumulator version with a dynamically indexed remainder went to local memory and ran 2.7x slower - reverted)
  double s1 = 0, s2 = 0;
  const float2* p = (const float2*)part + (size_t)b * nparts * C + c;
#pragma unroll 8
  for (int k = 0; k < nparts; k++) { const float2 v = __ldg(p + (size_t)k * C); s1 += v.x; s2 += v.y; }
  stats[(size_t)i * 2] = (float)s1; stats[(size_t)i * 2 + 1] = (float)s2;
}

// ------------------------------------------------------------------------------------------------ small layers
// first conv: x [B,4,H,W] f32 (NCHW, the agent's tensor), w [64][3][3][4] f32 -> y [B,H,W,64] bf16; no BN / ReLU follows (Modules.py:176)
__global__ void k_conv_first(const float* __restrict__ x, const float* __restrict__ w, bf16* __restrict__ y, int B, int H, int W) {
  __shared__ __align__(16) float sw[64 * 36];
  for (int i = threadIdx.x; i < 64 * 36; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  int p = blockIdx.x * blockDim.x + threadIdx.x, HW = H * W;
  if (p >= B * HW) return;
  int b = p / HW, oh = (p % HW) / W, ow = p % W;
  float in[36];
#pragma unroll
  for (int t = 0; t < 9; t++) {
    int ih = oh + t / 3 - 1, iw = ow + t % 3 - 1;
    bool ok = ih >= 0 && ih < H && iw >= 0 && iw < W;
#pragma unroll
    for (int c = 0; c < 4; c++) in[t * 4 + c] = ok ? x[((size_t)(b * 4 + c) * H + ih) * W + iw] : 0.f;
  }
  bf16* out = y + (size_t)p * 64;
  const float4* sw4 = (const float4*)sw;  // 9 x 16-byte broadcast reads per output channel instead of 36 scalar ones
  for (int co = 0; co < 64; co += 8) {
    __nv_bfloat162 o[4];
#pragma unroll
    for (int h = 0; h < 4; h++) {
      float a0 = 0, a1 = 0;
#pragma unroll
      for (int k = 0; k < 9; k++) {
        float4 w0 = sw4[(co + 2 * h) * 9 + k], w1 = sw4[(co + 2 * h + 1) * 9 + k];
        a0 += in[4 * k] * w0.x; a0 += in[4 * k + 1] * w0.y; a0 += in[4 * k + 2] * w0.z; a0 += in[4 * k + 3] * w0.w;
        a1 += in[4 * k] * w1.x; a1 += in[4 * k + 1] * w1.y; a1 += in[4 * k + 2] * w1.z; a1 += in[4 * k + 3] * w1.w;
      }
      o[h] = __floats2bfloat162_rn(a0, a1);
    }
    *(uint4*)(out + co) = *(uint4*)o;
  }
}
// same layer, two horizontally consecutive pixels per thread: every weight vector read from shared memory feeds both (half the LDS
// traffic per FMA, the limiter of the one-pixel version); pixel pairs are taken along the flattened index, B*H*W must be even
__global__ void __launch_bounds__(128) k_conv_first2(const float* __restrict__ x, const float* __restrict__ w, bf16* __restrict__ y, int B, int H, int W) {
  __shared__ __align__(16) float sw[64 * 36];
  for (int i = threadIdx.x; i < 64 * 36; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const int HW = H * W;
  const int p0 = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
  if (p0 >= B * HW) return;
  float in[2][36];
#pragma unroll
  for (int q = 0; q < 2; q++) {
    const int p = p0 + q, b = p / HW, oh = (p % HW) / W, ow = p % W;
#pragma unroll
    for (int t = 0; t < 9; t++) {
      const int ih = oh + t / 3 - 1, iw = ow + t % 3 - 1;
      const bool ok = ih >= 0 && ih < H && iw >= 0 && iw < W;
#pragma unroll
      for (int c = 0; c < 4; c++) in[q][t * 4 + c] = ok ? __ldg(x + ((size_t)(b * 4 + c) * H + ih) * W + iw) : 0.f;
    }
  }
  bf16* out = y + (size_t)p0 * 64;
  const float4* sw4 = (const float4*)sw;
  for (int co = 0; co < 64; co += 8) {
    __nv_bfloat162 o[2][4];
#pragma unroll
    for (int h = 0; h < 4; h++) {
      float a[2][2] = {{0, 0}, {0, 0}};
#pragma unroll
      for (int k = 0; k < 9; k++) {
        const float4 w0 = sw4[(co + 2 * h) * 9 + k], w1 = sw4[(co + 2 * h + 1) * 9 + k];
#pragma unroll
        for (int q = 0; q < 2; q++) {  // same accumulation order per output as k_conv_first
          a[q][0] += in[q][4 * k] * w0.x; a[q][0] += in[q][4 * k + 1] * w0.y; a[q][0] += in[q][4 * k + 2] * w0.z; a[q][0] += in[q][4 * k + 3] * w0.w;
          a[q][1] += in[q][4 * k] * w1.x; a[q][1] += in[q][4 * k + 1] * w1.y; a[q][1] += in[q][4 * k + 2] * w1.z; a[q][1] += in[q][4 * k + 3] * w1.w;
        }
      }
      o[0][h] = __floats2bfloat162_rn(a[0][0], a[0][1]); o[1][h] = __floats2bfloat162_rn(a[1][0], a[1][1]);
    }
    *(uint4*)(out + co) = *(uint4*)o[0];
    *(uint4*)(out + 64 + co) = *(uint4*)o[1];
  }
}
// MaxPool2d(kernel 3, stride 2, padding 1) on NHWC bf16
__global__ void k_maxpool(const bf16* __restrict__ x, bf16* __restrict__ y, int B, int H, int W, int C) {
  int OH = (H + 1) / 2, OW = (W + 1) / 2, vc = C / 8;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, n = (size_t)B * OH * OW * vc;
  if (i >= n) return;
  int cv = (i % vc) * 8;
  size_t p = i / vc;
  int ow = p % OW, oh = (p / OW) % OH, b = p / ((size_t)OW * OH);
  float m[8];
#pragma unroll
  for (int k = 0; k < 8; k++) m[k] = -3.0e38f;
  for (int dh = -1; dh <= 1; dh++) for (int dw = -1; dw <= 1; dw++) {
    int ih = 2 * oh + dh, iw = 2 * ow + dw;
    if (ih < 0 || ih >= H || iw < 0 || iw >= W) continue;
    uint4 qv = __ldg((const uint4*)(x + (((size_t)b * H + ih) * W + iw) * C + cv));
    const __nv_bfloat162* pv = (const __nv_bfloat162*)&qv;
#pragma unroll
    for (int k = 0; k < 4; k++) { float2 v = __bfloat1622float2(pv[k]); m[2 * k] = fmaxf(m[2 * k], v.x); m[2 * k + 1] = fmaxf(m[2 * k + 1], v.y); }
  }
  __nv_bfloat162 o[4];
#pragma unroll
  for (int k = 0; k < 4; k++) o[k] = __floats2bfloat162_rn(m[2 * k], m[2 * k + 1]);
  *(uint4*)(y + p * C + cv) = *(uint4*)o;
}
// y = relu( (x - mean) * rstd * gamma + beta [+ identity] ), per-image statistics from `stats` (sum, sumsq over H*W); fp32 in, bf16 out
// one CTA = a run of pixels of one image: per-channel scale / shift (training-mode BatchNorm with the image's own statistics) are
// prepared once in shared memory, then every thread streams 8 channels at a time (2 x 16 B in, 16 B out)
__global__ void __launch_bounds__(256) k_bn_act(const float* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ gamma,
                                                const float* __restrict__ beta, const float* __restrict__ identity, bf16* __restrict__ y, int HW, int C,
                                                float eps, int pix_per_cta) {
  extern __shared__ float sc[];  // scale[C], shift[C]
  const int b = blockIdx.y, tid = threadIdx.x;
  for (int c = tid; c < C; c += blockDim.x) {
    float s1 = stats[((size_t)b * C + c) * 2], s2 = stats[((size_t)b * C + c) * 2 + 1];
    float mean = s1 / HW, var = fmaxf(s2 / HW - mean * mean, 0.f);  // biased variance, as BatchNorm uses for normalisation
    float scale = rsqrtf(var + eps) * gamma[c];
    sc[c] = scale; sc[C + c] = beta[c] - mean * scale;
  }
  __syncthreads();
  const int vc = C / 8, p0 = blockIdx.x * pix_per_cta, np = min(HW, p0 + pix_per_cta) - p0;
  const size_t base = ((size_t)b * HW + p0) * C;
  for (int i = tid; i < np * vc; i += blockDim.x) {
    const int cv = (i % vc) * 8;
    const size_t off = base + (size_t)(i / vc) * C + cv;
    float v[8];
    *(float4*)v = __ldg((const float4*)(x + off)); *(float4*)(v + 4) = __ldg((const float4*)(x + off + 4));
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = v[k] * sc[cv + k] + sc[C + cv + k];
    if (identity) {
      float r[8];
      *(float4*)r = __ldg((const float4*)(identity + off)); *(float4*)(r + 4) = __ldg((const float4*)(identity + off + 4));
#pragma unroll
      for (int k = 0; k < 8; k++) v[k] += r[k];
    }
    __nv_bfloat162 o[4];
#pragma unroll
    for (int k = 0; k < 4; k++) o[k] = __floats2bfloat162_rn(fmaxf(v[2 * k], 0.f), fmaxf(v[2 * k + 1], 0.f));
    *(uint4*)(y + off) = *(uint4*)o;
  }
}
// nn.UpsamplingBilinear2d(scale_factor=2) = bilinear, align_corners=True; NHWC bf16, 8 channels (16 B) per thread
__global__ void k_upsample2x(const bf16* __restrict__ x, bf16* __restrict__ y, int B, int H, int W, int C) {
  int OH = 2 * H, OW = 2 * W, vc = C / 8;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, n = (size_t)B * OH * OW * vc;
  if (i >= n) return;
  int cv = (i % vc) * 8;
  size_t p = i / vc;
  int ow = p % OW, oh = (p / OW) % OH, b = p / ((size_t)OW * OH);
  float fy = oh * (float)(H - 1) / (float)(OH - 1), fx = ow * (float)(W - 1) / (float)(OW - 1);
  int y0 = (int)fy, x0 = (int)fx, y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
  float wy = fy - y0, wx = fx - x0;
  auto at = [&](int yy, int xx) { return __ldg((const uint4*)(x + (((size_t)b * H + yy) * W + xx) * C + cv)); };
  uint4 qa = at(y0, x0), qb = at(y0, x1), qc = at(y1, x0), qd = at(y1, x1);
  const __nv_bfloat162 *pa = (const __nv_bfloat162*)&qa, *pb = (const __nv_bfloat162*)&qb, *pc = (const __nv_bfloat162*)&qc, *pd = (const __nv_bfloat162*)&qd;
  __nv_bfloat162 o[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    float2 a = __bfloat1622float2(pa[k]), bb = __bfloat1622float2(pb[k]), cc = __bfloat1622float2(pc[k]), d = __bfloat1622float2(pd[k]);
    float r0 = (a.x * (1 - wx) + bb.x * wx) * (1 - wy) + (cc.x * (1 - wx) + d.x * wx) * wy;
    float r1 = (a.y * (1 - wx) + bb.y * wx) * (1 - wy) + (cc.y * (1 - wx) + d.y * wx) * wy;
    o[k] = __floats2bfloat162_rn(r0, r1);
  }
  *(uint4*)(y + p * C + cv) = *(uint4*)o;
}
// last layer: conv1x1 64 -> A (+bias) and sigmoid; x [B,HW,64] bf16, w [A][64] f32 -> q [B,A,HW] f32 (NCHW like the reference output)
// SIGMOID = false writes the pre-activation (used at low resolution by gq_head_up2)
template <bool SIGMOID>
__global__ void k_head(const bf16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ q, int B, int HW, int A) {
  __shared__ __align__(16) float sw[8 * 64 + 8];
  for (int i = threadIdx.x; i < A * 64; i += blockDim.x) sw[i] = w[i];
  if (threadIdx.x < A) sw[8 * 64 + threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= (size_t)B * HW) return;
  int b = p / HW, pix = p % HW;
  float in[64];
  const uint4* src = (const uint4*)(x + p * 64);  // the pixel's 64 channels = 8 x 16 bytes
#pragma unroll
  for (int c = 0; c < 8; c++) {
    const uint4 qv = __ldg(src + c);
    const __nv_bfloat162* pv = (const __nv_bfloat162*)&qv;
#pragma unroll
    for (int k = 0; k < 4; k++) { float2 v = __bfloat1622float2(pv[k]); in[8 * c + 2 * k] = v.x; in[8 * c + 2 * k + 1] = v.y; }
  }
  for (int a = 0; a < A; a++) {
    float acc = sw[8 * 64 + a];
#pragma unroll
    for (int c = 0; c < 16; c++) {
      float4 wv = ((const float4*)sw)[a * 16 + c];
      acc += in[4 * c] * wv.x; acc += in[4 * c + 1] * wv.y; acc += in[4 * c + 2] * wv.z; acc += in[4 * c + 3] * wv.w;
    }
    q[((size_t)b * A + a) * HW + pix] = SIGMOID ? 1.f / (1.f + __expf(-acc)) : acc;
  }
}
// q = sigmoid(UpsamplingBilinear2d(2)(z)) on A planes: z [B*A, H, W] f32 -> q [B*A, 2H, 2W] f32 (align_corners=True).
// The network ends with UP2 -> C1 (1x1 conv + bias) -> sigmoid (Modules.py:250-251,281-283); a 1x1 convolution is linear per pixel and the
// bilinear weights sum to one, so C1(UP2(x)) == UP2(C1(x)): the head runs on the 4x smaller map and only its A planes are up-sampled
__global__ void k_up2_sigmoid(const float* __restrict__ z, float* __restrict__ q, int planes, int H, int W) {
  const int OH = 2 * H, OW = 2 * W;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, n = (size_t)planes * OH * OW;
  if (i >= n) return;
  const int ow = i % OW, oh = (i / OW) % OH;
  const size_t pl = i / ((size_t)OW * OH);
  const float fy = oh * (float)(H - 1) / (float)(OH - 1), fx = ow * (float)(W - 1) / (float)(OW - 1);
  const int y0 = (int)fy, x0 = (int)fx, y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
  const float wy = fy - y0, wx = fx - x0;
  const float* zp = z + pl * H * W;
  const float a = __ldg(zp + y0 * W + x0), b = __ldg(zp + y0 * W + x1), c = __ldg(zp + y1 * W + x0), d = __ldg(zp + y1 * W + x1);
  const float v = (a * (1 - wx) + b * wx) * (1 - wy) + (c * (1 - wx) + d * wx) * wy;
  q[i] = 1.f / (1.f + __expf(-v));
}
// flat arg-max over the A*HW Q-values of every image (Grasping_Agent_multidiscrete.py:295-299): idx = rot*HW + y*W + x
__global__ void __launch_bounds__(1024) k_argmax(const float* __restrict__ q, int n, int* __restrict__ idx, float* __restrict__ val) {
  __shared__ float sv[1024];
  __shared__ int si[1024];
  const float* qb = q + (size_t)blockIdx.x * n;
  float bv = -1.f; int bi = 0;  // Q-values are sigmoids (> 0); ties go to the lowest flat index, like torch.max over the flattened map
  const int n4 = ((size_t)qb % 16 == 0) ? n / 4 : 0;
  for (int i = threadIdx.x; i < n4; i += blockDim.x) {
    float4 v = __ldg((const float4*)qb + i);
    if (v.x > bv) { bv = v.x; bi = 4 * i; }
    if (v.y > bv) { bv = v.y; bi = 4 * i + 1; }
    if (v.z > bv) { bv = v.z; bi = 4 * i + 2; }
    if (v.w > bv) { bv = v.w; bi = 4 * i + 3; }
  }
  for (int i = 4 * n4 + threadIdx.x; i < n; i += blockDim.x) { float v = qb[i]; if (v > bv) { bv = v; bi = i; } }
  sv[threadIdx.x] = bv; si[threadIdx.x] = bi;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) { float ov = sv[threadIdx.x + o]; int oi = si[threadIdx.x + o]; if (ov > sv[threadIdx.x] || (ov == sv[threadIdx.x] && oi < si[threadIdx.x])) { sv[threadIdx.x] = ov; si[threadIdx.x] = oi; } }
    __syncthreads();
  }
  if (threadIdx.x == 0) { idx[blockIdx.x] = si[0]; val[blockIdx.x] = sv[0]; }
}

// Grasp_Agent.transform_observation with normalize=True, jitter_and_noise=False (Grasping_Agent_multidiscrete.py:301-368), batched:
// depth clipped at `thr`, negated, min-max normalised per image; rgb u8 -> [0,1]; output state [B,4,H,W] f32 (r,g,b,depth)
__global__ void k_obs_minmax(const float* __restrict__ depth, int HW, float thr, float* __restrict__ mm) {
  __shared__ float smin[256], smax[256];
  const float* d = depth + (size_t)blockIdx.x * HW;
  float lo = 3.0e38f, hi = -3.0e38f;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) { float v = -fminf(d[i], thr); lo = fminf(lo, v); hi = fmaxf(hi, v); }
  smin[threadIdx.x] = lo; smax[threadIdx.x] = hi;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { smin[threadIdx.x] = fminf(smin[threadIdx.x], smin[threadIdx.x + o]); smax[threadIdx.x] = fmaxf(smax[threadIdx.x], smax[threadIdx.x + o]); }
    __syncthreads();
  }
  if (threadIdx.x == 0) { mm[2 * blockIdx.x] = smin[0]; mm[2 * blockIdx.x + 1] = smax[0]; }
}
__global__ void k_obs_state(const unsigned char* __restrict__ rgb, const float* __restrict__ depth, const float* __restrict__ mm, float thr, int B, int HW,
                            float* __restrict__ state) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * HW) return;
  int b = i / HW, p = i % HW;
  float lo = mm[2 * b], hi = mm[2 * b + 1];
  float* out = state + (size_t)b * 4 * HW + p;
  out[0] = rgb[3 * i] * (1.f / 255.f); out[HW] = rgb[3 * i + 1] * (1.f / 255.f); out[2 * HW] = rgb[3 * i + 2] * (1.f / 255.f);
  out[3 * (size_t)HW] = (-fminf(depth[i], thr) - lo) / (hi - lo);
}

// ---- training-time transform_observation (normalize=True, jitter_and_noise=True; Grasping_Agent_multidiscrete.py:118-124,301-368):
// depth clipped, + N(0, noise_std) per pixel, negated, min-max normalised per image AFTER the noise (:318-322); rgb through
// ColorJitter(brightness, contrast, saturation, hue) with the per-image factors and operation order drawn by the caller (torchvision's
// get_params: randperm(4) + four uniforms), applied with torchvision's float-tensor arithmetic (_blend / rgb_to_grayscale / _rgb2hsv /
// _hsv2rgb).  The noise comes from a counter-based generator (Philox4x32-10) keyed on the seed and indexed by (global env id, step,
// pixel): the value of a pixel does not depend on how the envs are batched or sharded over GPUs.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// standard normal from counter (pixel, env, step) under key = seed (Box-Muller on two of the four words)
__device__ __forceinline__ float obs_noise(uint64_t seed, uint64_t env, uint32_t step, uint32_t pixel) {
  uint32_t w[4];
  philox4x32_10(pixel, (uint32_t)env, (uint32_t)(env >> 32), step, (uint32_t)seed, (uint32_t)(seed >> 32), w);
  const float u1 = ((float)(w[0] >> 8) + 0.5f) * (1.0f / 16777216.0f), u2 = ((float)(w[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}
__device__ __forceinline__ float clamp01(float x) { return fminf(fmaxf(x, 0.f), 1.f); }
__device__ __forceinline__ float gray_of(float r, float g, float b) { return 0.2989f * r + 0.587f * g + 0.114f * b; }
// one ColorJitter operation on one pixel; op: 0 brightness, 1 contrast (needs the image's grey mean), 2 saturation, 3 hue
__device__ __forceinline__ void jitter_op(int op, float f, float mean, float& r, float& g, float& b) {
  if (op == 0) { r = clamp01(f * r); g = clamp01(f * g); b = clamp01(f * b); }
  else if (op == 1) { const float m = (1.f - f) * mean; r = clamp01(f * r + m); g = clamp01(f * g + m); b = clamp01(f * b + m); }
  else if (op == 2) { const float m = (1.f - f) * gray_of(r, g, b); r = clamp01(f * r + m); g = clamp01(f * g + m); b = clamp01(f * b + m); }
  else {
    const float maxc = fmaxf(r, fmaxf(g, b)), minc = fminf(r, fminf(g, b));
    const bool eqc = maxc == minc;
    const float cr = maxc - minc, s = cr / (eqc ? 1.f : maxc), div = eqc ? 1.f : cr;
    const float rc = (maxc - r) / div, gc = (maxc - g) / div, bc = (maxc - b) / div;
    const float hr = (maxc == r) ? (bc - gc) : 0.f, hg = (maxc == g && maxc != r) ? (2.f + rc - bc) : 0.f, hb = (maxc != g && maxc != r) ? (4.f + gc - rc) : 0.f;
    float h = fmodf((hr + hg + hb) / 6.f + 1.f, 1.f);
    h = h + f;                       // torch's % : result has the sign of the divisor
    h = h - floorf(h);
    const float v = maxc, h6 = h * 6.f, fl = floorf(h6), ff = h6 - fl;
    int i = ((int)fl) % 6;
    if (i < 0) i += 6;
    const float pp = clamp01(v * (1.f - s)), q = clamp01(v * (1.f - s * ff)), t = clamp01(v * (1.f - s * (1.f - ff)));
    r = i == 0 ? v : i == 1 ? q : i == 2 ? pp : i == 3 ? pp : i == 4 ? t : v;
    g = i == 0 ? t : i == 1 ? v : i == 2 ? v : i == 3 ? q : i == 4 ? pp : pp;
    b = i == 0 ? pp : i == 1 ? pp : i == 2 ? t : i == 3 ? v : i == 4 ? v : q;
  }
}
// per image: min / max of the noisy negated clipped depth, and the grey mean of the image as it is when the contrast operation runs
__global__ void __launch_bounds__(256) k_obs_train_reduce(const unsigned char* __restrict__ rgb, const float* __restrict__ depth, int HW, float thr, float noise_std,
                                                          unsigned long long seed, const long long* __restrict__ env_index, unsigned int step,
                                                          const float* __restrict__ jitter, const int* __restrict__ order, float* __restrict__ red) {
  __shared__ float smin[256], smax[256];
  __shared__ double ssum[256];
  const int b = blockIdx.x;
  const unsigned long long env = env_index ? (unsigned long long)env_index[b] : (unsigned long long)b;
  int ord[4] = {0, 1, 2, 3};
  float fac[4] = {1.f, 1.f, 1.f, 0.f};
  if (jitter) for (int k = 0; k < 4; k++) fac[k] = jitter[4 * b + k];
  if (order) for (int k = 0; k < 4; k++) ord[k] = order[4 * b + k];
  float lo = 3.0e38f, hi = -3.0e38f;
  double gs = 0;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) {
    float v = -(fminf(depth[(size_t)b * HW + i], thr) + noise_std * obs_noise(seed, env, step, (uint32_t)i));
    lo = fminf(lo, v); hi = fmaxf(hi, v);
    if (jitter) {
      const unsigned char* px = rgb + ((size_t)b * HW + i) * 3;
      float r = px[0] * (1.f / 255.f), g = px[1] * (1.f / 255.f), bl = px[2] * (1.f / 255.f);
      for (int k = 0; k < 4 && ord[k] != 1; k++) jitter_op(ord[k], fac[ord[k]], 0.f, r, g, bl);  // the operations in front of contrast
      gs += gray_of(r, g, bl);
    }
  }
  smin[threadIdx.x] = lo; smax[threadIdx.x] = hi; ssum[threadIdx.x] = gs;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      smin[threadIdx.x] = fminf(smin[threadIdx.x], smin[threadIdx.x + o]); smax[threadIdx.x] = fmaxf(smax[threadIdx.x], smax[threadIdx.x + o]);
      ssum[threadIdx.x] += ssum[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { red[4 * b] = smin[0]; red[4 * b + 1] = smax[0]; red[4 * b + 2] = (float)(ssum[0] / HW); red[4 * b + 3] = 0.f; }
}
__global__ void k_obs_train_state(const unsigned char* __restrict__ rgb, const float* __restrict__ depth, const float* __restrict__ red, float thr, float noise_std,
                                  unsigned long long seed, const long long* __restrict__ env_index, unsigned int step, const float* __restrict__ jitter,
                                  const int* __restrict__ order, int B, int HW, float* __restrict__ state, bf16* __restrict__ state_nhwc) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * HW) return;
  const int b = i / HW, p = i % HW;
  const unsigned long long env = env_index ? (unsigned long long)env_index[b] : (unsigned long long)b;
  float r = rgb[3 * i] * (1.f / 255.f), g = rgb[3 * i + 1] * (1.f / 255.f), bl = rgb[3 * i + 2] * (1.f / 255.f);
  if (jitter) {
    const float mean = red[4 * b + 2];
    for (int k = 0; k < 4; k++) { const int op = order ? order[4 * b + k] : k; jitter_op(op, jitter[4 * b + op], mean, r, g, bl); }
  }
  const float lo = red[4 * b], hi = red[4 * b + 1];
  const float d = (-(fminf(depth[i], thr) + noise_std * obs_noise(seed, env, step, (uint32_t)p)) - lo) / (hi - lo);
  if (state) {
    float* out = state + (size_t)b * 4 * HW + p;
    out[0] = r; out[HW] = g; out[2 * (size_t)HW] = bl; out[3 * (size_t)HW] = d;
  }
  if (state_nhwc) {
    __nv_bfloat162* o = (__nv_bfloat162*)(state_nhwc + i * 4);
    o[0] = __floats2bfloat162_rn(r, g); o[1] = __floats2bfloat162_rn(bl, d);
  }
}

// ------------------------------------------------------------------------------------------------ C-ABI
extern "C" int gq_obs_to_state_train(const unsigned char* rgb, const float* depth, float depth_threshold, float noise_std, unsigned long long seed,
                                     const long long* env_index, unsigned int step, const float* jitter, const int* order, float* scratch_red,
                                     float* state, void* state_nhwc_bf16, int B, int HW, void* stream) {
  if (!rgb || !depth || !scratch_red || (!state && !state_nhwc_bf16) || B <= 0 || HW <= 0) { snprintf(q_err, sizeof q_err, "gq_obs_to_state_train: bad argument"); return -1; }
  cudaStream_t st = (cudaStream_t)stream;
  k_obs_train_reduce<<<B, 256, 0, st>>>(rgb, depth, HW, depth_threshold, noise_std, seed, env_index, step, jitter, order, scratch_red);
  k_obs_train_state<<<(unsigned)(((size_t)B * HW + 255) / 256), 256, 0, st>>>(rgb, depth, scratch_red, depth_threshold, noise_std, seed, env_index, step, jitter, order,
                                                                              B, HW, state, (bf16*)state_nhwc_bf16);
  QCK(cudaGetLastError());
  return 0;
}
extern "C" int gq_obs_to_state(const unsigned char* rgb, const float* depth, float depth_threshold, float* scratch_minmax, float* state, int B, int HW,
                               void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  k_obs_minmax<<<B, 256, 0, st>>>(depth, HW, depth_threshold, scratch_minmax);
  k_obs_state<<<(unsigned)(((size_t)B * HW + 255) / 256), 256, 0, st>>>(rgb, depth, scratch_minmax, depth_threshold, B, HW, state);
  QCK(cudaGetLastError());
  return 0;
}
extern "C" const char* gq_version(void) { return "grasp_qnet 0.5 sm_100a bf16 tcgen05 (persistent CTAs, both operands by TMA: im2col activations + tiled weights, SW128 K-major, double-buffered TMEM accumulator, fused block tail)"; }

// tile shape: 128 output pixels x 64 / 128 / 256 output channels (the widest that divides Cout).  Tuning overrides: GQ_BN=128,
// GQ_PERSIST=0 (non-persistent kernel: 3 stages, GQ_NPW=4|8 producer warps, GQ_CGB=1 weights bypass L1, GQ_TMA=0 weight tiles by cp.async).
struct ConvCfg { int bn, npw, cg, tma, persist; };
static ConvCfg conv_cfg(int Cout) {
  static int env_bn = -1, env_npw = 8, env_cg = 0, env_tma = 1, env_persist = 2;
  if (env_bn < 0) {
    const char* e = getenv("GQ_BN"); env_bn = e ? atoi(e) : 0;
    e = getenv("GQ_NPW"); env_npw = (e && atoi(e) == 4) ? 4 : 8;  // r01i sweep, whole forward: 4 warps 366, 8 warps 381 TFLOP/s
    e = getenv("GQ_CGB"); env_cg = (e && atoi(e) != 0) ? 1 : 0;
    e = getenv("GQ_PERSIST"); env_persist = (e && atoi(e) == 0) ? 0 : 2;  // r02a: persistent TMA kernel 477, non-persistent 439 TFLOP/s
    e = getenv("GQ_TMA"); env_tma = (e && atoi(e) == 0) ? 0 : 1;  // r01k: weight tiles by TMA 439 vs 414 TFLOP/s whole forward
  }
  ConvCfg c;
  c.bn = (Cout % 128 == 0) ? 128 : 64;
  if (env_bn != 128 && Cout % 256 == 0) c.bn = 256;
  c.npw = env_npw; c.cg = env_cg; c.tma = env_tma; c.persist = env_persist;
  return c;
}
static size_t conv_smem(int bn, int nst) { return (size_t)nst * (BM * BK * 2) + (size_t)nst * ((size_t)bn * BK * 2) + 8 * (2 * nst + 1) + 16; }

// 2-D tensor map of the packed weights w [Cout][taps * Cin] bf16 with a [BLOCK_N rows x 64 columns] box, 128-byte swizzle (the layout the
// UMMA descriptors expect); cuTensorMapEncodeTiled comes from the driver through the runtime (no libcuda link dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int make_weight_tmap(CUtensorMap* m, const void* w, int Cout, int K, int bn) {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
  }
  if (!fn) { snprintf(q_err, sizeof q_err, "cuTensorMapEncodeTiled is not available"); return -3; }
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)Cout};
  cuuint64_t gstride[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)bn};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { snprintf(q_err, sizeof q_err, "cuTensorMapEncodeTiled failed (%d)", (int)r); return -3; }
  return 0;
}

// im2col-mode tensor map of the activations x [N][H][W][C] bf16: 64 channels x 128 output positions per load, zero fill outside the image
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const int*, const int*,
                                   cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                   CUtensorMapFloatOOBfill);
static int make_act_tmap_im2col(CUtensorMap* m, const void* x, int B, int H, int W, int C, int ks) {
  static EncodeIm2colFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess) fn = (EncodeIm2colFn)p;
  }
  if (!fn) { snprintf(q_err, sizeof q_err, "cuTensorMapEncodeIm2col is not available"); return -3; }
  const int pad = ks / 2;
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t gstride[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  int lower[2] = {-pad, -pad}, upper[2] = {pad - (ks - 1), pad - (ks - 1)};  // fprop corners (W, H): as many box positions as output pixels
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), gdim, gstride, lower, upper, (cuuint32_t)BK, (cuuint32_t)BM, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { snprintf(q_err, sizeof q_err, "cuTensorMapEncodeIm2col failed (%d)", (int)r); return -3; }
  return 0;
}

// EPI 0: y (+ BN partials); EPI 1: out = relu(resid * scale + shift + conv + bias) bf16
template <int EPI>
static int launch_conv(const ConvCfg& c, dim3 grid, cudaStream_t st, const bf16* x, const bf16* w, const float* bias, float* y, float* partials,
                       const float* resid, const float2* ss, bf16* out, int H, int W, int Cin, int Cout, int ks) {
  const size_t smem = conv_smem(c.bn, 3);
  alignas(64) CUtensorMap tm;
  memset(&tm, 0, sizeof tm);
  bool tma = c.tma != 0;
  if (tma && make_weight_tmap(&tm, w, Cout, ks * ks * Cin, c.bn) != 0) {  // no encoder in this driver: same kernel with cp.async weight copies
    static bool told = false;
    if (!told) { told = true; fprintf(stderr, "grasp_qnet: %s; weight tiles fall back to cp.async\n", q_err); }
    tma = false;
  }
  if (c.persist && tma) {
    // one CTA per SM, static tile list, 192 KB of stages (4 / 6 / 8 at BLOCK_N 256 / 128 / 64)
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int gx = (int)grid.x, gy = (int)grid.y, ntiles = gx * gy * (int)grid.z;
    const int nctas = ntiles < sms ? ntiles : sms;
    alignas(64) CUtensorMap tmx;
    static bool im2col_ok = true;
    if (im2col_ok && make_act_tmap_im2col(&tmx, x, (int)grid.z, H, W, Cin, ks) != 0) {
      im2col_ok = false;
      fprintf(stderr, "grasp_qnet: %s; using the non-persistent kernel\n", q_err);
    }
    if (im2col_ok) {
#define LAUNCH_T(BN_, ST_)                                                                                                           \
  do {                                                                                                                               \
    const size_t psmem = (size_t)ST_ * (BM * BK * 2 + BN_ * BK * 2) + 8 * (2 * ST_ + 4) + 16;                                         \
    QCK(cudaFuncSetAttribute(k_conv_tc_tma<BN_, ST_, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));                 \
    k_conv_tc_tma<BN_, ST_, EPI><<<nctas, 192, psmem, st>>>(tm, tmx, bias, y, partials, resid, ss, out, H, W, Cin, Cout, ks, gx, gy, ntiles); \
  } while (0)
      if (c.bn == 256) LAUNCH_T(256, 4);
      else if (c.bn == 128) LAUNCH_T(128, 6);
      else LAUNCH_T(64, 8);
#undef LAUNCH_T
      return 0;
    }
  }
#define LAUNCH_ONE(BN_, NPW_, TMA_)                                                                                                  \
  do {                                                                                                                               \
    QCK(cudaFuncSetAttribute(k_conv_tc<BN_, 3, EPI, NPW_, TMA_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));            \
    k_conv_tc<BN_, 3, EPI, NPW_, TMA_><<<grid, 32 * (NPW_ + 1), smem, st>>>(tm, x, w, bias, y, partials, resid, ss, out, H, W, Cin, Cout, ks, c.cg); \
  } while (0)
#define LAUNCH_CONV(BN_)                                                                                                             \
  do {                                                                                                                               \
    if (c.npw == 8) { if (tma) LAUNCH_ONE(BN_, 8, 1); else LAUNCH_ONE(BN_, 8, 0); }                                                  \
    else { if (tma) LAUNCH_ONE(BN_, 4, 1); else LAUNCH_ONE(BN_, 4, 0); }                                                             \
  } while (0)
  if (c.bn == 256) LAUNCH_CONV(256);
  else if (c.bn == 128) LAUNCH_CONV(128);
  else LAUNCH_CONV(64);
#undef LAUNCH_CONV
#undef LAUNCH_ONE
  return 0;
}

extern "C" int gq_conv_tc(const void* x, const void* w, const float* bias, float* y, float* stats, float* partials, int B, int H, int W, int Cin, int Cout, int ks,
                          void* stream) {
  if (!x || !w || !y || (ks != 1 && ks != 3) || Cin % 64 || Cout % 64) { snprintf(q_err, sizeof q_err, "gq_conv_tc: bad argument"); return -1; }
  if (stats && !partials) { snprintf(q_err, sizeof q_err, "gq_conv_tc: stats requested without a partials buffer"); return -1; }
  cudaStream_t st = (cudaStream_t)stream;
  const ConvCfg c = conv_cfg(Cout);
  dim3 grid((H * W + BM - 1) / BM, Cout / c.bn, B);
  int r = launch_conv<0>(c, grid, st, (const bf16*)x, (const bf16*)w, bias, y, stats ? partials : nullptr, nullptr, nullptr, nullptr, H, W, Cin, Cout, ks);
  if (r) return r;
  if (stats) k_bn_reduce<<<(B * Cout + 127) / 128, 128, 0, st>>>(partials, stats, B, (int)grid.x * 4, Cout);
  QCK(cudaGetLastError());
  return 0;
}
extern "C" int gq_conv_tc_block_out(const void* x, const void* w, const float* bias, const float* resid, const float* stats, const float* gamma,
                                    const float* beta, float eps, float* scratch_scale_shift, void* out, int B, int H, int W, int Cin, int Cout, int ks,
                                    void* stream) {
  if (!x || !w || !resid || !stats || !gamma || !beta || !scratch_scale_shift || !out || (ks != 1 && ks != 3) || Cin % 64 || Cout % 64) {
    snprintf(q_err, sizeof q_err, "gq_conv_tc_block_out: bad argument");
    return -1;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const ConvCfg c = conv_cfg(Cout);
  dim3 grid((H * W + BM - 1) / BM, Cout / c.bn, B);
  k_bn_scale_shift<<<(B * Cout + 127) / 128, 128, 0, st>>>(stats, gamma, beta, (float2*)scratch_scale_shift, B, Cout, H * W, eps);
  int r = launch_conv<1>(c, grid, st, (const bf16*)x, (const bf16*)w, bias, nullptr, nullptr, resid, (const float2*)scratch_scale_shift, (bf16*)out, H, W, Cin, Cout, ks);
  if (r) return r;
  QCK(cudaGetLastError());
  return 0;
}
extern "C" int gq_conv_first(const float* x, const float* w, void* y, int B, int H, int W, void* stream) {
  static int one_pixel = -1;
  if (one_pixel < 0) { const char* e = getenv("GQ_FIRST1"); one_pixel = (e && atoi(e) != 0) ? 1 : 0; }  // GQ_FIRST1=1: one pixel per thread
  if (one_pixel || ((size_t)B * H * W) % 2) k_conv_first<<<(B * H * W + 127) / 128, 128, 0, (cudaStream_t)stream>>>(x, w, (bf16*)y, B, H, W);
  else k_conv_first2<<<(B * H * W / 2 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(x, w, (bf16*)y, B, H, W);
  QCK(cudaGetLastError());
  return 0;
}
extern "C" int gq_maxpool(const void* x, void* y, int B, int H, int W, int C, void* stream) {
  size_t n = (size_t)B * ((H + 1) / 2) * ((W + 1) / 2) * (C / 8);
  k_maxpool<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)y, B, H, W, C);
  QCK(cudaGetLastError());
  return 0;
}
extern "C" int gq_bn_act(const float* x, const float* stats, const float* gamma, const float* beta, const float* identity, void* y, int B, int HW, int C,
                         float eps, void* stream) {
  if (C % 8) { snprintf(q_err, sizeof q_err, "gq_bn_act: C must be a multiple of 8"); return -1; }
  // enough CTAs per image to fill the GPU (148 SMs x 8) without making the per-CTA prologue (2 C floats) matter
  int ctas = (148 * 8 + B - 1) / B;
  int ppc = (HW + ctas - 1) / ctas;
  if (ppc < 16) ppc = 16;
  dim3 grid((HW + ppc - 1) / ppc, B);
  k_bn_act<<<grid, 256, 2 * C * sizeof(float), (cudaStream_t)stream>>>(x, stats, gamma, beta, identity, (bf16*)y, HW, C, eps, ppc);
  QCK(cudaGetLastError());
  return 0;
}
extern "C" int gq_upsample2x(const void* x, void* y, int B, int H, int W, int C, void* stream) {
  if (C % 8) { snprintf(q_err, sizeof q_err, "gq_upsample2x: C must be a multiple of 8"); return -1; }
  size_t n = (size_t)B * 4 * H * W * (C / 8);
  k_upsample2x<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)y, B, H, W, C);
  QCK(cudaGetLastError());
  return 0;
}
extern "C" int gq_head(const void* x, const float* w, const float* bias, float* q, int B, int HW, int A, void* stream) {
  if (A > 8) { snprintf(q_err, sizeof q_err, "gq_head: at most 8 action channels"); return -1; }
  k_head<true><<<(unsigned)(((size_t)B * HW + 127) / 128), 128, 0, (cudaStream_t)stream>>>((const bf16*)x, w, bias, q, B, HW, A);
  QCK(cudaGetLastError());
  return 0;
}
extern "C" int gq_head_up2(const void* x, const float* w, const float* bias, float* scratch_z, float* q, int B, int H, int W, int A, void* stream) {
  if (A > 8 || !scratch_z) { snprintf(q_err, sizeof q_err, "gq_head_up2: bad argument"); return -1; }
  cudaStream_t st = (cudaStream_t)stream;
  k_head<false><<<(unsigned)(((size_t)B * H * W + 127) / 128), 128, 0, st>>>((const bf16*)x, w, bias, scratch_z, B, H * W, A);
  size_t n = (size_t)B * A * 4 * H * W;
  k_up2_sigmoid<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(scratch_z, q, B * A, H, W);
  QCK(cudaGetLastError());
  return 0;
}
extern "C" int gq_argmax(const float* q, int B, int n, int* idx, float* val, void* stream) {
  k_argmax<<<B, 1024, 0, (cudaStream_t)stream>>>(q, n, idx, val);
  QCK(cudaGetLastError());
  return 0;
}
#include "qnet_learn.cuh"
