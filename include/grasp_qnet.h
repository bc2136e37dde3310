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
#include <stddef.h>
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

/* ------------------------------------------------------------------------------------------------------------------------------------
 * Learner: the compute of Grasp_Agent.learn() (Grasping_Agent_multidiscrete.py:388-446, optimiser :153-156) - see csrc/qnet_learn.cuh.
 * All buffers [dev]; activations NHWC bf16, pre-BatchNorm outputs fp32, gradients of activations bf16, of parameters fp32. */
/* BatchNorm over the whole batch (learn() forwards B images at once, training mode): per-image sums -> batch sums, in place.
 * stats [B,C,2] as written by gq_conv_tc. */
int gq_bn_batch_merge(float* stats, int B, int C, void* stream);
/* q_pred = net(state).view(B,-1).gather(1, action); loss = binary_cross_entropy(q_pred, reward) (Grasping_Agent_multidiscrete.py:426-443)
 * + backward through sigmoid, the bilinear 2x up-sampling of the head planes and the 1x1 head convolution (Modules.py:250-251,281-283):
 * q [B,A,2H,2W] f32 (forward output), x [B,H,W,64] bf16 (head input), w_head [A,64] f32, action [B] int64 (rot*4HW + y*2W + x), reward [B] f32.
 * Out: loss_terms [B] (mean = the loss), qsel [B] = q_pred, dX [B,H,W,64] bf16, dW_head [A,64], db_head [A]; partials [B,A,64], [B,A]. */
int gq_loss_head_bwd(const float* q, const void* x, const float* w_head, const long long* action, const float* reward, int B, int A, int H, int W,
                     float* loss_terms, float* qsel, void* dX, float* dW_partial, float* db_partial, float* dW_head, float* db_head, void* stream);
/* backward of y = relu(BatchNorm(o) [+ identity]) with batch statistics (BasicBlock, Modules.py:128-142): dY, act = y [B,HW,C] bf16, o [B,HW,C] f32,
 * stats merged [B,C,2], gamma [C].  Out: d_o [B,HW,C] bf16 (gradient w.r.t. the convolution output), dgamma, dbeta [C] (either may be NULL),
 * dpre_out (NULL or [B,HW,C] bf16) = dY * (y > 0), the gradient that also enters the identity branch.
 * scratch_part: ceil(B*HW/128) * C * 2 floats, scratch_sums: 2 C floats. */
int gq_bn_relu_bwd(const void* dY, const void* act, const float* o, const float* stats, const float* gamma, int B, int HW, int C, float eps,
                   void* dpre_out, float* scratch_part, float* scratch_sums, float* dgamma, float* dbeta, void* d_o, void* stream);
/* weight gradient of conv3x3 / conv1x1 (Modules.py:145-156): dY [B,H,W,Cout] bf16, x [B,H,W,Cin] bf16 -> dW [Cout][ks*ks][Cin] f32;
 * scratch_part: B * Cout * ks*ks * Cin floats (split-K over the images, reduced in image order) */
int gq_conv_wgrad(const void* dY, const void* x, float* scratch_part, float* dW, int B, int H, int W, int Cin, int Cout, int ks, void* stream);
/* weight gradient of Perception_Module.C1 (4 -> 64, Modules.py:163): dY [B,H,W,64] bf16, x [B,4,H,W] f32 -> dW [64][3][3][4] f32;
 * scratch_part: ceil(B*H*W/2048) * 64 * 36 floats */
int gq_conv_first_wgrad(const void* dY, const float* x, float* scratch_part, float* dW, int B, int H, int W, void* stream);
/* MaxPool2d(3, 2, 1) backward: x [B,H,W,C] bf16 (pool input), dY [B,ceil(H/2),ceil(W/2),C] bf16 -> dX [B,H,W,C] bf16 */
int gq_maxpool_bwd(const void* x, const void* dY, void* dX, int B, int H, int W, int C, void* stream);
/* UpsamplingBilinear2d(2) backward: dY [B,2H,2W,C] bf16 -> dX [B,H,W,C] bf16 */
int gq_upsample2x_bwd(const void* dY, void* dX, int B, int H, int W, int C, void* stream);
/* out (bf16) = a (f32) + b (f32, may be NULL) or a + c_bf16 (bf16, may be NULL): sums two gradient branches */
int gq_add_to_bf16(const float* a, const float* b, const void* c_bf16, void* out, size_t n, void* stream);
/* torch.optim.Adam step with weight_decay as L2 (Grasping_Agent_multidiscrete.py:153-156): p, g, m, v [n] f32, step counted from 1 */
int gq_adam(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1, float beta2, float eps, float weight_decay, int step, void* stream);

#ifdef __cplusplus
}
#endif
#endif
