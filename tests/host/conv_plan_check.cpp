// Host check of mujoco_rl_ur5_b200/csrc/qnet_plan.h (the hoisted gather arithmetic of the convolution producers) against the direct
// per-copy definition: for every layer shape of the Q-network and the odd shapes of the GPU parity tests, every M tile, every thread,
// every k-step and every row the plan must name the same global source, the same shared-memory destination, the same zero-fill
// decision, the same stage and the same mbarrier wait parity.  Exit code 0 = identical.
#include <cstdio>
#include <cstdlib>
#include <initializer_list>

#include "../../mujoco_rl_ur5_b200/csrc/qnet_plan.h"

struct Shape { int H, W, Cin, Cout, ks; };

template <int RS>
static long check(const Shape& s, int BLOCK_N, int STAGES) {
  const int HW = s.H * s.W, pad = s.ks / 2, taps = s.ks * s.ks, kchunks = s.Cin / QP_BK, nk = taps * kchunks;
  const int A_STAGE = QP_BM * QP_BK * 2, B_STAGE = BLOCK_N * QP_BK * 2;
  long n = 0;
  for (int m0 = 0; m0 < HW; m0 += QP_BM)
    for (int n0 = 0; n0 < s.Cout; n0 += BLOCK_N)
      for (int tid = 0; tid < 8 * RS; tid++) {
        ConvPlan p;
        conv_plan_init<RS>(p, tid, m0, n0, s.H, s.W, s.Cin, s.ks);
        const int chunk = tid & 7, rbase = tid >> 3;
        for (int kn = 0; kn < nk; kn++) {
          // ---- definition (first kernel version)
          const int sn = kn % STAGES, tap = kn / kchunks, c0 = (kn % kchunks) * QP_BK, dh = tap / s.ks - pad, dw = tap % s.ks - pad;
          const bool wait = kn >= STAGES;
          const unsigned parity = wait ? (unsigned)(((kn - STAGES) / STAGES) & 1) : 0u;
          if (p.sn != sn || (p.round > 0) != wait || (wait && (((unsigned)(p.round - 1)) & 1u) != parity) || p.tap != tap) {
            printf("state mismatch: kn %d sn %d/%d round %d tap %d/%d\n", kn, p.sn, sn, p.round, p.tap, tap);
            return -1;
          }
          const ptrdiff_t a0 = conv_plan_a(p, s.W, s.Cin);
          for (int i = 0; i < QP_BM / RS; i++) {
            const int row = rbase + RS * i, mm = m0 + row;
            const int oh = mm / s.W, ow = mm % s.W, ih = oh + dh, iw = ow + dw;
            const bool ok = mm < HW && ih >= 0 && ih < s.H && iw >= 0 && iw < s.W;
            const ptrdiff_t src = (((ptrdiff_t)ih * s.W + iw) * s.Cin + c0 + chunk * 8) * 2;
            const unsigned dst = (unsigned)(sn * A_STAGE + row * 128 + ((chunk ^ (row & 7)) << 4));
            const bool pok = (p.vmask[i] >> p.tap) & 1u;
            const unsigned pdst = (unsigned)(p.sn * A_STAGE) + p.dstoff + (unsigned)i * (RS * 128u);
            if (pok != ok || pdst != dst || (ok && a0 + i * p.a_stride != src)) {
              printf("A mismatch: shape %dx%d cin %d ks %d m0 %d tid %d kn %d i %d ok %d/%d dst %u/%u src %td/%td\n", s.H, s.W, s.Cin, s.ks, m0, tid, kn, i,
                     (int)pok, (int)ok, pdst, dst, a0 + i * p.a_stride, src);
              return -1;
            }
            if (ok && (src < 0 || src + 16 > (ptrdiff_t)HW * s.Cin * 2)) { printf("A source out of the image\n"); return -1; }
            n++;
          }
          const ptrdiff_t b0 = conv_plan_b(p, kn);
          for (int i = 0; i < BLOCK_N / RS; i++) {
            const int row = rbase + RS * i;
            const ptrdiff_t src = ((((ptrdiff_t)(n0 + row)) * taps + tap) * s.Cin + c0 + chunk * 8) * 2;
            const unsigned dst = (unsigned)(sn * B_STAGE + row * 128 + ((chunk ^ (row & 7)) << 4));
            const unsigned pdst = (unsigned)(p.sn * B_STAGE) + p.dstoff + (unsigned)i * (RS * 128u);
            if (pdst != dst || b0 + i * p.b_stride != src || src + 16 > (ptrdiff_t)s.Cout * taps * s.Cin * 2) {
              printf("B mismatch: cin %d cout %d ks %d n0 %d tid %d kn %d i %d dst %u/%u src %td/%td\n", s.Cin, s.Cout, s.ks, n0, tid, kn, i, pdst, dst,
                     b0 + i * p.b_stride, src);
              return -1;
            }
            n++;
          }
          conv_plan_next(p, STAGES, kchunks, pad);
        }
      }
  return n;
}

int main() {
  const Shape shapes[] = {
      // the network (Modules.py:159-287): 0.RB1 @100x100, 0.RB2 / 0.RB3 / 1.RB1 / 1.RB2 @50x50, 1.RB3 @100x100 (3x3, 3x3, 1x1 each)
      {100, 100, 64, 128, 3}, {100, 100, 128, 128, 3}, {100, 100, 64, 128, 1}, {50, 50, 128, 256, 3}, {50, 50, 256, 256, 3}, {50, 50, 128, 256, 1},
      {50, 50, 256, 512, 3}, {50, 50, 512, 512, 3}, {50, 50, 256, 512, 1}, {50, 50, 512, 256, 3}, {50, 50, 512, 256, 1}, {50, 50, 256, 128, 3},
      {50, 50, 128, 128, 3}, {50, 50, 256, 128, 1}, {100, 100, 128, 64, 3}, {100, 100, 64, 64, 3}, {100, 100, 128, 64, 1},
      // odd shapes: partial last tile, width not a multiple of anything, tiny images (rows wrap several times inside 16 pixels)
      {37, 41, 64, 64, 3}, {37, 41, 64, 64, 1}, {5, 7, 64, 64, 3}, {1, 1, 64, 64, 3}, {3, 200, 64, 128, 3}, {200, 3, 128, 64, 3}, {16, 16, 64, 256, 3}};
  long total = 0;
  for (const Shape& s : shapes)
    for (int bn : {64, 128, 256}) {
      if (s.Cout % bn) continue;
      for (int st : {3, 4}) {
        long n = check<16>(s, bn, st), n2 = check<32>(s, bn, st);  // 4 and 8 producer warps
        if (n < 0 || n2 < 0) return 1;
        total += n + n2;
      }
    }
  printf("conv plan: %ld copies identical to the definition\n", total);
  return 0;
}
