/* grasp_qnet.h — C-ABI of the batched pixel-wise grasp Q-network forward (libgrasp_qnet.so).
 *
 * Stands in for `policy_net(state)` in the reference agent (Grasping_Agent_multidiscrete.py:254,295) where
 * policy_net = Modules.MULTIDISCRETE_RESNET(6) (Modules.py:308-311).  The layer functions below are the building blocks
 * (one per nn.Module type the network uses); mujoco_rl_ur5_b200/qnet.py strings them together in the order of
 * Perception_Module.forward (Modules.py:170-193) and Grasping_Module_multidiscrete.forward (:256-287).
 * All pointers are CUDA device pointers; activations are NHWC bf16 unless stated; every call is asynchronous on `stream`
 * (cudaStream_t cast to void*).  Return 0 on success, negative on error (gq_last_error()).
 */
#ifndef GRASP_QNET_H
#define GRASP_QNET_H
#ifdef __cplusplus
extern "C" {
#endif

const char* gq_last_error(void);
const char* gq_version(void);

/* nn.Conv2d(Cin, Cout, kernel_size=ks, padding=ks/2, stride=1) for ks in {1,3}, Cin and Cout multiples of 64 (conv3x3: Modules.py:145-156;
 * BasicBlock.conv3 1x1 with bias: :126).  tcgen05 implicit GEMM.  x [B,H,W,Cin] bf16, w [Cout][ks*ks][Cin] bf16, bias [Cout] f32 or NULL,
 * y [B,H,W,Cout] f32.  If stats != NULL, stats[B][Cout][2] f32 := per-image (sum, sum of squares) of y over H*W, computed
 * deterministically through `partials`, a scratch buffer of B * ceil(H*W/128) * 4 * Cout * 2 floats. */
int gq_conv_tc(const void* x, const void* w, const float* bias, float* y, float* stats, float* partials, int B, int H, int W, int Cin, int Cout, int ks,
               void* stream);
/* Tail of BasicBlock.forward (Modules.py:136-142: `out = bn2(conv2(out)); identity = conv3(x); out += identity; out = relu(out)`) with the
 * shortcut convolution doing the rest in its epilogue: out = relu((resid - mean) * rsqrt(var + eps) * gamma + beta + conv(x, w) + bias) as
 * bf16 [B,H,W,Cout].  resid [B,H,W,Cout] f32 and stats [B,Cout,2] are conv2's outputs from gq_conv_tc; gamma / beta [Cout] are bn2's;
 * scratch_scale_shift: B * Cout * 2 floats.  Same shapes and constraints as gq_conv_tc. */
int gq_conv_tc_block_out(const void* x, const void* w, const float* bias, const float* resid, const float* stats, const float* gamma, const float* beta,
                         float eps, float* scratch_scale_shift, void* out, int B, int H, int W, int Cin, int Cout, int ks, void* stream);
/* Perception_Module.C1 = conv3x3(4, 64), no bias (Modules.py:163): x [B,4,H,W] f32 NCHW, w [64][3][3][4] f32, y [B,H,W,64] bf16 */
int gq_conv_first(const float* x, const float* w, void* y, int B, int H, int W, void* stream);
/* nn.MaxPool2d(kernel_size=3, stride=2, padding=1) (Modules.py:164,166): x [B,H,W,C] -> y [B,ceil(H/2),ceil(W/2),C] */
int gq_maxpool(const void* x, void* y, int B, int H, int W, int C, void* stream);
/* nn.BatchNorm2d in training mode with per-image statistics + optional residual add + ReLU (BasicBlock.forward, Modules.py:128-142):
 * y = relu((x - mean) * rsqrt(var + eps) * gamma + beta [+ identity]); x, identity [B,HW,C] f32, stats from gq_conv_tc, y bf16 */
int gq_bn_act(const float* x, const float* stats, const float* gamma, const float* beta, const float* identity, void* y, int B, int HW, int C,
              float eps, void* stream);
/* nn.UpsamplingBilinear2d(scale_factor=2) (align_corners=True) (Modules.py:248,250): x [B,H,W,C] -> y [B,2H,2W,C] */
int gq_upsample2x(const void* x, void* y, int B, int H, int W, int C, void* stream);
/* Grasping_Module_multidiscrete.C1 = nn.Conv2d(64, A, 1) + squeeze + Sigmoid (Modules.py:251,281-283): x [B,HW,64] bf16, w [A][64] f32,
 * bias [A] f32 -> q [B,A,HW] f32 (the reference's NCHW output layout) */
int gq_head(const void* x, const float* w, const float* bias, float* q, int B, int HW, int A, void* stream);
/* The same tail computed at the lower resolution: Grasping_Module_multidiscrete ends with UP2 -> C1 -> sigmoid (Modules.py:250-251,281-283) and
 * C1 (1x1 conv + bias) commutes with bilinear up-sampling, so q = sigmoid(UP2(C1(x))): x [B,H,W,64] bf16 (the input of UP2),
 * scratch_z [B,A,H,W] f32, q [B,A,2H,2W] f32 */
int gq_head_up2(const void* x, const float* w, const float* bias, float* scratch_z, float* q, int B, int H, int W, int A, void* stream);
/* Grasp_Agent.transform_observation(normalize=True, jitter_and_noise=False) batched (Grasping_Agent_multidiscrete.py:301-368):
 * rgb [B,HW,3] u8, depth [B,HW] f32 metres -> state [B,4,HW] f32 (rgb/255, depth clipped at depth_threshold, negated, min-max per image);
 * scratch_minmax [B,2] f32 */
int gq_obs_to_state(const unsigned char* rgb, const float* depth, float depth_threshold, float* scratch_minmax, float* state, int B, int HW,
                    void* stream);
/* Grasp_Agent.transform_observation(normalize=True, jitter_and_noise=True) batched - what the agent feeds the network while it learns
 * (Grasping_Agent_multidiscrete.py:118-124 ColorJitter(0.5, 0.5, 0.5, 0.5), :318 depth += N(0, 0.001) before the min-max normalisation).
 * noise_std: standard deviation of the depth noise (0.001); seed / env_index [B] int64 (NULL: 0..B-1) / step: key and counter of the
 * counter-based generator, so a pixel's noise depends on (seed, global env id, step, pixel) only; jitter [B,4] f32 = brightness,
 * contrast, saturation, hue factors and order [B,4] int32 = the permutation of the four operations, both drawn by the caller as
 * torchvision's ColorJitter.get_params does (NULL jitter: no colour change; NULL order: 0,1,2,3); scratch_red [B,4] f32.
 * Outputs (either may be NULL): state [B,4,HW] f32 (the agent's tensor layout), state_nhwc_bf16 [B,HW,4] bf16 (NHWC, feeds the first
 * convolution without a repack). */
int gq_obs_to_state_train(const unsigned char* rgb, const float* depth, float depth_threshold, float noise_std, unsigned long long seed,
                          const long long* env_index, unsigned int step, const float* jitter, const int* order, float* scratch_red, float* state,
                          void* state_nhwc_bf16, int B, int HW, void* stream);
/* output.view(-1).max(0) per image (Grasping_Agent_multidiscrete.py:295-299): q [B,n] -> idx [B] int32 (rot*HW + y*W + x), val [B] f32 */
int gq_argmax(const float* q, int B, int n, int* idx, float* val, void* stream);

#ifdef __cplusplus
}
#endif
#endif
