// Small fp64 device helpers for the warp-per-environment physics kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define GE_MINVAL 1e-15
#define FULL 0xffffffffu

namespace ge {

__device__ __forceinline__ void v3set(double* r, double a, double b, double c) { r[0] = a; r[1] = b; r[2] = c; }
__device__ __forceinline__ void v3copy(double* r, const double* a) { r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; }
__device__ __forceinline__ void v3add(double* r, const double* a, const double* b) { r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2]; }
__device__ __forceinline__ void v3sub(double* r, const double* a, const double* b) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
__device__ __forceinline__ void v3scl(double* r, const double* a, double s) { r[0] = a[0] * s; r[1] = a[1] * s; r[2] = a[2] * s; }
__device__ __forceinline__ void v3addscl(double* r, const double* a, const double* b, double s) { r[0] = a[0] + b[0] * s; r[1] = a[1] + b[1] * s; r[2] = a[2] + b[2] * s; }
__device__ __forceinline__ double v3dot(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ __forceinline__ void v3cross(double* r, const double* a, const double* b) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
__device__ __forceinline__ double v3norm(const double* a) { return sqrt(v3dot(a, a)); }
__device__ __forceinline__ double v3normalize(double* a) {
  double n = v3norm(a);
  if (n < GE_MINVAL) { a[0] = 1; a[1] = 0; a[2] = 0; return 0; }
  double inv = 1.0 / n;
  a[0] *= inv; a[1] *= inv; a[2] *= inv;
  return n;
}
// r = M(row-major 3x3) * v
__device__ __forceinline__ void m3mulv(double* r, const double* m, const double* v) {
  double x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2], y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2], z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
__device__ __forceinline__ void m3Tmulv(double* r, const double* m, const double* v) {
  double x = m[0] * v[0] + m[3] * v[1] + m[6] * v[2], y = m[1] * v[0] + m[4] * v[1] + m[7] * v[2], z = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
__device__ __forceinline__ void m3mul(double* r, const double* a, const double* b) {
  double t[9];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) t[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
#pragma unroll
  for (int i = 0; i < 9; i++) r[i] = t[i];
}
__device__ __forceinline__ void m3col(double* r, const double* m, int k) { r[0] = m[k]; r[1] = m[3 + k]; r[2] = m[6 + k]; }
__device__ __forceinline__ void qmul(double* r, const double* a, const double* b) {
  double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  double x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  double y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  double z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
__device__ __forceinline__ void qnormalize(double* q) {
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < GE_MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  double inv = 1.0 / n;
  q[0] *= inv; q[1] *= inv; q[2] *= inv; q[3] *= inv;
}
__device__ __forceinline__ void qaxisangle(double* q, const double* axis, double angle) {
  double s, c;
  sincos(0.5 * angle, &s, &c);
  q[0] = c; q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}
__device__ __forceinline__ void q2mat(double* m, const double* q) {
  double w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = w * w + x * x - y * y - z * z; m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
  m[3] = 2 * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2 * (y * z - w * x);
  m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}
// rotate vector by unit quaternion
__device__ __forceinline__ void qrot(double* r, const double* q, const double* v) {
  double m[9];
  q2mat(m, q);
  m3mulv(r, m, v);
}
__device__ __forceinline__ double dot6(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5]; }
// f(6) = I(10) * v(6); I = [m, h(3), Ixx Iyy Izz Ixy Ixz Iyz] about the world origin
__device__ __forceinline__ void inert_mul(double* f, const double* I, const double* v) {
  const double *w = v, *l = v + 3, *h = I + 1;
  double hv[3], wh[3];
  v3cross(hv, h, l); v3cross(wh, w, h);
  f[0] = I[4] * w[0] + I[7] * w[1] + I[8] * w[2] + hv[0];
  f[1] = I[7] * w[0] + I[5] * w[1] + I[9] * w[2] + hv[1];
  f[2] = I[8] * w[0] + I[9] * w[1] + I[6] * w[2] + hv[2];
  f[3] = I[0] * l[0] + wh[0]; f[4] = I[0] * l[1] + wh[1]; f[5] = I[0] * l[2] + wh[2];
}
__device__ __forceinline__ void cross_motion(double* r, const double* v, const double* s) {
  double a[3], b[3], c[3];
  v3cross(a, v, s); v3cross(b, v, s + 3); v3cross(c, v + 3, s);
  v3copy(r, a); v3add(r + 3, b, c);
}
__device__ __forceinline__ void cross_force(double* r, const double* v, const double* f) {
  double a[3], b[3], c[3];
  v3cross(a, v, f); v3cross(b, v + 3, f + 3); v3cross(c, v, f + 3);
  v3add(r, a, b); v3copy(r + 3, c);
}
// warp reductions (fixed butterfly order => deterministic)
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
  return v;
}
// ---- primitives over the GE_LANES threads that share one environment (ge_variant.h): a warp, or a CTA of GE_NW warps
#ifndef GE_LANES
#define GE_LANES 32
#define GE_NW 1
#endif
#if GE_NW == 1
__device__ __forceinline__ void gsync() { __syncwarp(); }
__device__ __forceinline__ double group_sum(double v) { return warp_sum(v); }
__device__ __forceinline__ double group_max(double v) { return warp_max(v); }
__device__ __forceinline__ bool group_any(bool p) { return __any_sync(FULL, p) != 0; }
__device__ __forceinline__ int group_bcast_int(int v, int lane) { return __shfl_sync(FULL, v, 0); }
// exclusive prefix of one int per thread in thread order, and the total (ordered compaction)
__device__ __forceinline__ int group_exscan(int v, int lane, int& total) {
  int off = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(FULL, off, o); if (lane >= o) off += u; }
  total = __shfl_sync(FULL, off, 31);
  return off - v;
}
#else
// __syncthreads() is an ALIGNED barrier: every warp has to arrive converged.  The code in front of a group barrier often ends in
// branches on values that are equal across the lanes only mathematically (MPR iterations, Newton / line-search exits), after which
// the hardware may still run the lanes of a warp separately - r02d: "illegal instruction" at >= 160 envs, found by synccheck.  So every
// group barrier first reconverges its warp.
__device__ __forceinline__ void cta_sync() { __syncwarp(); __syncthreads(); }
__device__ __forceinline__ void gsync() { cta_sync(); }
// fixed order: butterfly inside each warp, then warp 0 .. GE_NW-1 (deterministic)
__device__ __forceinline__ double group_sum(double v) {
  __shared__ double red[GE_NW];
  __syncwarp();
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  cta_sync();
  double s = red[0];
#pragma unroll
  for (int w = 1; w < GE_NW; w++) s += red[w];
  cta_sync();
  return s;
}
__device__ __forceinline__ double group_max(double v) {
  __shared__ double red[GE_NW];
  __syncwarp();
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  cta_sync();
  double s = red[0];
#pragma unroll
  for (int w = 1; w < GE_NW; w++) s = fmax(s, red[w]);
  cta_sync();
  return s;
}
// (a flag in shared memory and plain barriers instead of __syncthreads_or: r02d's GPU core dump showed "Warp Illegal Instruction" at the
// BAR.RED of this function with the other warps of the CTA still leaving the preceding BAR.SYNC - the two barrier kinds share hardware
// barrier 0 and PTX forbids intermixing bar.red with bar.sync on an active barrier)
__device__ __forceinline__ bool group_any(bool p) {
  __shared__ int flag;
  if (threadIdx.x == 0) flag = 0;
  cta_sync();
  if (p) flag = 1;
  cta_sync();
  const bool r = flag != 0;
  cta_sync();
  return r;
}
__device__ __forceinline__ int group_bcast_int(int v, int lane) {
  __shared__ int b;
  if (lane == 0) b = v;
  cta_sync();
  int r = b;
  cta_sync();
  return r;
}
__device__ __forceinline__ int group_exscan(int v, int lane, int& total) {
  __shared__ int wsum[GE_NW];
  const int wl = lane & 31, wid = lane >> 5;
  int off = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(FULL, off, o); if (wl >= o) off += u; }
  if (wl == 31) wsum[wid] = off;
  cta_sync();
  int before = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < GE_NW; w++) { if (w < wid) before += wsum[w]; tot += wsum[w]; }
  cta_sync();
  total = tot;
  return before + off - v;
}
#endif

// argmax with lowest-index tie break; returns the winning (value, index) on every lane
__device__ __forceinline__ void warp_argmax(double& v, int& idx) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double ov = __shfl_xor_sync(FULL, v, o);
    int oi = __shfl_xor_sync(FULL, idx, o);
    if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
}
}  // namespace ge
