// Gather plan of the convolution producers (k_conv_tc in qnet.cu): which 16-byte chunks a producer thread copies in k-step kn,
// from where in global memory (activations x [HW][Cin] of one image, weights w [Cout][taps][Cin]) to where in the stage's
// 128-byte-swizzled K-major shared-memory tiles.  Plain integer code shared by the kernel and by a host check
// (tests/host/conv_plan_check.cpp compares it, for every thread / k-step / row, with the direct per-copy formulas of the first
// kernel version), so that the hoisted arithmetic cannot drift from the definition.
#pragma once
#include <stddef.h>
#include <stdint.h>
#ifndef __CUDACC__
#define QP_HD inline
#else
#define QP_HD __host__ __device__ __forceinline__
#endif

#define QP_BM 128  // output pixels per tile
#define QP_BK 64   // input channels per k-step (one 128-byte swizzle row of bf16)

struct ConvPlan {
  // constant per thread
  ptrdiff_t a_row0;   // byte offset from the image's activations of (row rbase, tap (0,0), channel chunk) ; row i adds i * a_stride
  ptrdiff_t a_stride; // RS pixels
  ptrdiff_t b_row0;   // byte offset from the weights of (output channel n0 + rbase, k-step 0) ; row i adds i * b_stride, k-step kn adds kn * 128
  ptrdiff_t b_stride; // RS output channels
  uint32_t dstoff;    // byte offset inside a stage tile of (row rbase, this thread's swizzled chunk); row i adds i * RS * 128
  uint32_t vmask[8];  // bit t: tap t of A row i reads inside the image (else the copy zero-fills)
  // running state
  int sn, round, tap, kc, dh, dw;
};

// RS = row step = producer threads / 8 (16 with 4 producer warps, 32 with 8): thread tid in [0, 8 RS) copies the 16-byte chunk
// tid % 8 of rows tid / 8 + RS i, i < 128 / RS (A) or BLOCK_N / RS (B).  ks is 1 or 3.
template <int RS>
QP_HD void conv_plan_init(ConvPlan& p, int tid, int m0, int n0, int H, int W, int Cin, int ks) {
  const int chunk = tid & 7, rbase = tid >> 3, taps = ks * ks, HW = H * W, pad = ks / 2;
  p.dstoff = (uint32_t)(rbase * 128 + ((chunk ^ (rbase & 7)) << 4));  // (rbase + RS i) % 8 == rbase % 8: one swizzle term per thread
  p.a_row0 = ((ptrdiff_t)(m0 + rbase) * Cin + chunk * 8) * 2;
  p.a_stride = (ptrdiff_t)RS * Cin * 2;
  p.b_row0 = ((ptrdiff_t)(n0 + rbase) * taps * Cin + chunk * 8) * 2;
  p.b_stride = (ptrdiff_t)RS * taps * Cin * 2;
  int oh = (m0 + rbase) / W, ow = (m0 + rbase) - oh * W;
#pragma unroll
  for (int i = 0; i < QP_BM / RS; i++) {
    uint32_t mk = 0;
    if (m0 + rbase + RS * i < HW) {
      if (ks == 1) mk = 1u;
      else {
        const uint32_t wb = (ow >= 1 ? 1u : 0u) | 2u | (ow + 1 < W ? 4u : 0u);
        mk = (oh >= 1 ? wb : 0u) | (wb << 3) | (oh + 1 < H ? wb << 6 : 0u);
      }
    }
    p.vmask[i] = mk;
    ow += RS;
    while (ow >= W) { ow -= W; oh++; }
  }
  p.sn = 0; p.round = 0; p.tap = 0; p.kc = 0; p.dh = -pad; p.dw = -pad;
}
// byte offset (from the image's activations) of A row 0 of the current k-step; row i adds i * a_stride
QP_HD ptrdiff_t conv_plan_a(const ConvPlan& p, int W, int Cin) { return p.a_row0 + ((ptrdiff_t)(p.dh * W + p.dw) * Cin + p.kc * QP_BK) * 2; }
// byte offset (from the weights) of B row 0 of k-step kn
QP_HD ptrdiff_t conv_plan_b(const ConvPlan& p, int kn) { return p.b_row0 + (ptrdiff_t)kn * (QP_BK * 2); }
// advance to the next k-step
QP_HD void conv_plan_next(ConvPlan& p, int stages, int kchunks, int pad) {
  if (++p.sn == stages) { p.sn = 0; p.round++; }
  if (++p.kc == kchunks) { p.kc = 0; p.tap++; if (++p.dw > pad) { p.dw = -pad; p.dh++; } }
}
