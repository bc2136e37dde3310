"""Pixel-wise grasp Q-network: weights container + batched bf16 forward on the in-tree CUDA library (libgrasp_qnet.so).

Reference: Modules.py:308-311 `MULTIDISCRETE_RESNET(number_actions_dim_2)` = nn.Sequential(Perception_Module (:159-193),
Grasping_Module_multidiscrete (:243-287)), built from BasicBlock (:92-142) and conv3x3 (:145-156).

* `make_torch_qnet()` re-declares the architecture with the reference's attribute names and construction order, so that
  (a) `torch.manual_seed(s)` gives bit-identical initial weights to the reference module and (b) the reference's checkpoints
  (`model_state_dict`, Grasping_Agent_multidiscrete.py:560-572) load with `load_state_dict`.  It is a weights container and the
  fp32 comparison target of the tests — the product forward never calls it.
* `QNetForward` packs the weights once (bf16, [Cout][kh][kw][Cin]) and runs the whole forward with the hand-written kernels:
  tcgen05 implicit-GEMM convolutions + fused BN/ReLU/residual, max-pool, bilinear up-sampling, sigmoid head, arg-max.
  BatchNorm uses per-image batch statistics: the reference never calls `.eval()` on the policy net and forwards one image at a
  time, so a batch of N environments is N independent batch-1 forwards.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
QLIB_PATH = os.path.join(_HERE, "csrc", "libgrasp_qnet.so")
_QLIB = None
QSYMBOLS = ["gq_last_error", "gq_version", "gq_conv_tc", "gq_conv_tc_block_out", "gq_conv_first", "gq_maxpool", "gq_bn_act", "gq_upsample2x", "gq_head", "gq_head_up2", "gq_argmax", "gq_obs_to_state", "gq_obs_to_state_train"]


def load_qnet_library():
    global _QLIB
    if _QLIB is None:
        if not os.path.exists(QLIB_PATH):
            raise RuntimeError(f"{QLIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(QLIB_PATH)
        P, I, F = C.c_void_p, C.c_int, C.c_float
        L.gq_last_error.restype = C.c_char_p
        L.gq_version.restype = C.c_char_p
        L.gq_conv_tc.argtypes = [P, P, P, P, P, P, I, I, I, I, I, I, P]
        L.gq_conv_tc_block_out.argtypes = [P, P, P, P, P, P, P, F, P, P, I, I, I, I, I, I, P]
        L.gq_conv_first.argtypes = [P, P, P, I, I, I, P]
        L.gq_maxpool.argtypes = [P, P, I, I, I, I, P]
        L.gq_bn_act.argtypes = [P, P, P, P, P, P, I, I, I, F, P]
        L.gq_upsample2x.argtypes = [P, P, I, I, I, I, P]
        L.gq_head.argtypes = [P, P, P, P, I, I, I, P]
        L.gq_head_up2.argtypes = [P, P, P, P, P, I, I, I, I, P]
        L.gq_argmax.argtypes = [P, I, I, P, P, P]
        L.gq_obs_to_state.argtypes = [P, P, F, P, P, I, I, P]
        L.gq_obs_to_state_train.argtypes = [P, P, F, F, C.c_uint64, P, C.c_uint32, P, P, P, P, P, I, I, P]
        _QLIB = L
    return _QLIB


def make_torch_qnet(number_actions_dim_2=6):
    """torch.nn re-declaration of MULTIDISCRETE_RESNET with the reference's parameter names and creation order."""
    import torch.nn as nn

    def conv3x3(i, o):
        return nn.Conv2d(i, o, kernel_size=3, stride=1, padding=1, bias=False)

    class BasicBlock(nn.Module):
        def __init__(self, inplanes, planes):
            super().__init__()
            self.conv1 = conv3x3(inplanes, planes)
            self.bn1 = nn.BatchNorm2d(planes)
            self.relu = nn.ReLU(inplace=True)
            self.conv2 = conv3x3(planes, planes)
            self.bn2 = nn.BatchNorm2d(planes)
            self.conv3 = nn.Conv2d(inplanes, planes, kernel_size=1, stride=1) if inplanes != planes else None

        def forward(self, x):
            identity = x
            out = self.relu(self.bn1(self.conv1(x)))
            out = self.bn2(self.conv2(out))
            if self.conv3 is not None:
                identity = self.conv3(identity)
            return self.relu(out + identity)

    class Perception_Module(nn.Module):
        def __init__(self):
            super().__init__()
            self.C1 = conv3x3(4, 64)
            self.MP1 = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
            self.RB1 = BasicBlock(64, 128)
            self.MP2 = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
            self.RB2 = BasicBlock(128, 256)
            self.RB3 = BasicBlock(256, 512)

        def forward(self, x):
            return self.RB3(self.RB2(self.MP2(self.RB1(self.MP1(self.C1(x))))))

    class Grasping_Module_multidiscrete(nn.Module):
        def __init__(self, act_dim_2):
            super().__init__()
            self.RB1 = BasicBlock(512, 256)
            self.RB2 = BasicBlock(256, 128)
            self.UP1 = nn.UpsamplingBilinear2d(scale_factor=2)
            self.RB3 = BasicBlock(128, 64)
            self.UP2 = nn.UpsamplingBilinear2d(scale_factor=2)
            self.C1 = nn.Conv2d(64, act_dim_2, kernel_size=1)
            self.sigmoid = nn.Sigmoid()

        def forward(self, x):
            x = self.C1(self.UP2(self.RB3(self.UP1(self.RB2(self.RB1(x))))))
            return self.sigmoid(x.squeeze())

    return nn.Sequential(Perception_Module(), Grasping_Module_multidiscrete(number_actions_dim_2))


_BLOCKS = ["0.RB1", "0.RB2", "0.RB3", "1.RB1", "1.RB2", "1.RB3"]


class QNetForward:
    """Batched forward on cuda.  `state_dict` uses the reference's key names ('0.C1.weight', '0.RB1.conv1.weight', ...)."""

    def __init__(self, state_dict, device=0, max_batch=64):
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("QNetForward needs a CUDA device (there is no CPU fallback)")
        self.torch = torch
        self.L = load_qnet_library()
        self.dev = torch.device("cuda", device)
        self.max_batch = max_batch
        sd = {k: v.detach().to(self.dev, torch.float32) for k, v in state_dict.items()}
        self.A = sd["1.C1.weight"].shape[0]
        pack = lambda w: w.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)  # [Cout,Cin,kh,kw] -> [Cout,kh,kw,Cin]
        self.w_first = sd["0.C1.weight"].permute(0, 2, 3, 1).contiguous()       # [64,3,3,4] f32
        self.blocks = []
        for name in _BLOCKS:
            b = dict(w1=pack(sd[f"{name}.conv1.weight"]), w2=pack(sd[f"{name}.conv2.weight"]), w3=pack(sd[f"{name}.conv3.weight"]),
                     b3=sd[f"{name}.conv3.bias"].contiguous(), g1=sd[f"{name}.bn1.weight"].contiguous(), be1=sd[f"{name}.bn1.bias"].contiguous(),
                     g2=sd[f"{name}.bn2.weight"].contiguous(), be2=sd[f"{name}.bn2.bias"].contiguous(),
                     cin=sd[f"{name}.conv1.weight"].shape[1], cout=sd[f"{name}.conv1.weight"].shape[0])
            self.blocks.append(b)
        # BasicBlock tail: shortcut conv + BN-apply of the main branch + add + ReLU in one kernel (GQ_FUSE_TAIL=0: separate kernels)
        self.fuse_tail = os.environ.get("GQ_FUSE_TAIL", "1") != "0"
        # network tail: head on the 100x100 map, then up-sample its 6 planes (GQ_FUSE_HEAD=0: up-sample 64 channels, then the head)
        self.fuse_head = os.environ.get("GQ_FUSE_HEAD", "1") != "0"
        self.w_head = sd["1.C1.weight"].reshape(self.A, 64).contiguous()
        self.b_head = sd["1.C1.bias"].contiguous()
        self.launches = 0
        # GQ_GRAPH=1: the ~110 launches of one chunk captured once per input shape in a CUDA graph and replayed (static input / output
        # buffers; tensor maps and pointers are baked into the graph).  Off by default: see DESIGN.md for the measurement.
        self.use_graph = os.environ.get("GQ_GRAPH", "0") == "1"
        self._graphs = {}

    def _ck(self, r, what):
        if r != 0:
            raise RuntimeError(f"{what} failed ({r}): {self.L.gq_last_error().decode()}")
        self.launches += 1

    def _p(self, t):
        return None if t is None else C.c_void_p(t.data_ptr())

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.dev).cuda_stream)

    # ---- layer wrappers (also used one by one by the parity tests)
    def conv_tc(self, x, w, bias, B, H, W, cin, cout, ks, want_stats):
        t = self.torch
        y = t.empty((B, H, W, cout), dtype=t.float32, device=self.dev)
        stats = t.empty((B, cout, 2), dtype=t.float32, device=self.dev) if want_stats else None
        part = t.empty((B, (H * W + 127) // 128 * 4, cout, 2), dtype=t.float32, device=self.dev) if want_stats else None
        self._ck(self.L.gq_conv_tc(self._p(x), self._p(w), self._p(bias), self._p(y), self._p(stats), self._p(part), B, H, W, cin, cout, ks, self._stream()), "gq_conv_tc")
        return y, stats

    def conv_tc_block_out(self, x, w, bias, resid, stats, gamma, beta, B, H, W, cin, cout, ks=1):
        """relu(BN(resid; per-image stats, gamma, beta) + conv(x, w) + bias) -> bf16 [B, H*W, cout]: the tail of BasicBlock.forward
        (Modules.py:136-142) with the shortcut convolution's epilogue doing the normalisation, the add and the ReLU"""
        t = self.torch
        y = t.empty((B, H * W, cout), dtype=t.bfloat16, device=self.dev)
        ss = t.empty((B, cout, 2), dtype=t.float32, device=self.dev)
        self._ck(self.L.gq_conv_tc_block_out(self._p(x), self._p(w), self._p(bias), self._p(resid), self._p(stats), self._p(gamma), self._p(beta), 1e-5,
                                              self._p(ss), self._p(y), B, H, W, cin, cout, ks, self._stream()), "gq_conv_tc_block_out")
        return y

    def bn_act(self, x, stats, gamma, beta, identity, B, HW, Cc):
        t = self.torch
        y = t.empty((B, HW, Cc), dtype=t.bfloat16, device=self.dev)
        self._ck(self.L.gq_bn_act(self._p(x), self._p(stats), self._p(gamma), self._p(beta), self._p(identity), self._p(y), B, HW, Cc, 1e-5, self._stream()), "gq_bn_act")
        return y

    def basic_block(self, x, blk, B, H, W):
        """BasicBlock.forward (Modules.py:128-142) on NHWC bf16 `x`"""
        cin, cout = blk["cin"], blk["cout"]
        o1, s1 = self.conv_tc(x, blk["w1"], None, B, H, W, cin, cout, 3, True)
        a1 = self.bn_act(o1, s1, blk["g1"], blk["be1"], None, B, H * W, cout)
        o2, s2 = self.conv_tc(a1, blk["w2"], None, B, H, W, cout, cout, 3, True)
        if self.fuse_tail:
            return self.conv_tc_block_out(x, blk["w3"], blk["b3"], o2, s2, blk["g2"], blk["be2"], B, H, W, cin, cout, 1)
        idn, _ = self.conv_tc(x, blk["w3"], blk["b3"], B, H, W, cin, cout, 1, False)
        return self.bn_act(o2, s2, blk["g2"], blk["be2"], idn, B, H * W, cout)

    def maxpool(self, x, B, H, W, Cc):
        t = self.torch
        y = t.empty((B, (H + 1) // 2, (W + 1) // 2, Cc), dtype=t.bfloat16, device=self.dev)
        self._ck(self.L.gq_maxpool(self._p(x), self._p(y), B, H, W, Cc, self._stream()), "gq_maxpool")
        return y

    def upsample(self, x, B, H, W, Cc):
        t = self.torch
        y = t.empty((B, 2 * H, 2 * W, Cc), dtype=t.bfloat16, device=self.dev)
        self._ck(self.L.gq_upsample2x(self._p(x), self._p(y), B, H, W, Cc, self._stream()), "gq_upsample2x")
        return y

    def _forward_chunk(self, state):
        t = self.torch
        B, _, H, W = state.shape
        x = t.empty((B, H, W, 64), dtype=t.bfloat16, device=self.dev)
        self._ck(self.L.gq_conv_first(self._p(state), self._p(self.w_first), self._p(x), B, H, W, self._stream()), "gq_conv_first")
        x = self.maxpool(x, B, H, W, 64)
        h, w = (H + 1) // 2, (W + 1) // 2
        x = self.basic_block(x, self.blocks[0], B, h, w)            # 0.RB1 64->128 @100
        x = self.maxpool(x, B, h, w, 128)
        h2, w2 = (h + 1) // 2, (w + 1) // 2
        x = self.basic_block(x, self.blocks[1], B, h2, w2)          # 0.RB2 128->256 @50
        x = self.basic_block(x, self.blocks[2], B, h2, w2)          # 0.RB3 256->512
        x = self.basic_block(x, self.blocks[3], B, h2, w2)          # 1.RB1 512->256
        x = self.basic_block(x, self.blocks[4], B, h2, w2)          # 1.RB2 256->128
        x = self.upsample(x.view(B, h2, w2, 128), B, h2, w2, 128)   # -> 100
        x = self.basic_block(x, self.blocks[5], B, 2 * h2, 2 * w2)  # 1.RB3 128->64
        q = t.empty((B, self.A, 4 * h2, 4 * w2), dtype=t.float32, device=self.dev)
        if self.fuse_head:
            z = t.empty((B, self.A, 2 * h2, 2 * w2), dtype=t.float32, device=self.dev)
            self._ck(self.L.gq_head_up2(self._p(x), self._p(self.w_head), self._p(self.b_head), self._p(z), self._p(q), B, 2 * h2, 2 * w2, self.A, self._stream()),
                     "gq_head_up2")
            return q
        x = self.upsample(x.view(B, 2 * h2, 2 * w2, 64), B, 2 * h2, 2 * w2, 64)  # -> 200
        HW = 16 * h2 * w2
        self._ck(self.L.gq_head(self._p(x), self._p(self.w_head), self._p(self.b_head), self._p(q), B, HW, self.A, self._stream()), "gq_head")
        return q

    def forward(self, state):
        """state [B,4,H,W] f32 on the device -> Q [B,A,H,W] f32 (sigmoid), chunked to bound activation memory"""
        t = self.torch
        state = state.to(self.dev, t.float32).contiguous()
        fn = self._forward_chunk_graphed if self.use_graph else self._forward_chunk
        outs = [fn(state[i:i + self.max_batch]) for i in range(0, state.shape[0], self.max_batch)]
        return outs[0] if len(outs) == 1 else t.cat(outs, dim=0)

    def _forward_chunk_graphed(self, state):
        t = self.torch
        key = tuple(state.shape)
        g = self._graphs.get(key)
        if g is None:
            sin = state.clone()
            cur = t.cuda.current_stream(self.dev)
            side = t.cuda.Stream(self.dev)
            side.wait_stream(cur)
            with t.cuda.stream(side):  # warm-up outside the capture (function attributes, driver entry points, allocator)
                self._forward_chunk(sin)
            cur.wait_stream(side)
            n0 = self.launches
            graph = t.cuda.CUDAGraph()
            with t.cuda.graph(graph):
                sout = self._forward_chunk(sin)
            g = self._graphs[key] = (graph, sin, sout, self.launches - n0)
        graph, sin, sout, nl = g
        sin.copy_(state)
        graph.replay()
        self.launches += nl
        return sout.clone()

    def obs_to_state(self, obs, depth_threshold=1.1):
        """batched deterministic transform_observation: obs dict of device tensors (rgb u8 [B,H,W,3], depth f32 [B,H,W]) -> [B,4,H,W] f32.
        depth_threshold = cam height - TABLE_HEIGHT + 0.01 = 1.1 (Grasping_Agent_multidiscrete.py:130-135)"""
        t = self.torch
        rgb, depth = obs["rgb"].contiguous(), obs["depth"].contiguous()
        B, H, W = depth.shape
        state = t.empty((B, 4, H, W), dtype=t.float32, device=self.dev)
        mm = t.empty((B, 2), dtype=t.float32, device=self.dev)
        self._ck(self.L.gq_obs_to_state(self._p(rgb), self._p(depth), float(depth_threshold), self._p(mm), self._p(state), B, H * W, self._stream()), "gq_obs_to_state")
        self.launches += 1
        return state

    def draw_color_jitter(self, B, generator=None, brightness=0.5, contrast=0.5, saturation=0.5, hue=0.5):
        """per-image ColorJitter parameters exactly as torchvision's ColorJitter.get_params draws them (randperm(4), then the four
        uniforms, per image; Grasping_Agent_multidiscrete.py:121): -> (factors [B,4] f32, order [B,4] int32) on the device"""
        t = self.torch
        fac, order = [], []
        for _ in range(B):
            order.append(t.randperm(4, generator=generator))
            fac.append(t.stack([t.empty(1).uniform_(1 - brightness, 1 + brightness, generator=generator)[0],
                                t.empty(1).uniform_(1 - contrast, 1 + contrast, generator=generator)[0],
                                t.empty(1).uniform_(1 - saturation, 1 + saturation, generator=generator)[0],
                                t.empty(1).uniform_(-hue, hue, generator=generator)[0]]))
        return t.stack(fac).to(self.dev, t.float32).contiguous(), t.stack(order).to(self.dev, t.int32).contiguous()

    def obs_to_state_train(self, obs, depth_threshold=1.1, noise_std=0.001, seed=0, step=0, env_index=None, jitter=None, order=None, nhwc_bf16=False):
        """transform_observation as the LEARNING agent calls it (normalize=True, jitter_and_noise=True), batched on the device: depth noise
        N(0, noise_std) from a counter-based generator keyed on (seed, global env id, step, pixel) added before the min-max normalisation,
        ColorJitter with the given per-image factors / operation order (see draw_color_jitter; None = no colour change).
        -> [B,4,H,W] f32, or ([B,4,H,W] f32, [B,H,W,4] bf16 NHWC) with nhwc_bf16=True"""
        t = self.torch
        rgb, depth = obs["rgb"].contiguous(), obs["depth"].contiguous()
        B, H, W = depth.shape
        state = t.empty((B, 4, H, W), dtype=t.float32, device=self.dev)
        nhwc = t.empty((B, H, W, 4), dtype=t.bfloat16, device=self.dev) if nhwc_bf16 else None
        red = t.empty((B, 4), dtype=t.float32, device=self.dev)
        ei = None if env_index is None else t.as_tensor(env_index).to(self.dev, t.int64).contiguous()
        self._ck(self.L.gq_obs_to_state_train(self._p(rgb), self._p(depth), float(depth_threshold), float(noise_std), int(seed) & (2 ** 64 - 1), self._p(ei),
                                              int(step) & 0xFFFFFFFF, self._p(jitter), self._p(order), self._p(red), self._p(state), self._p(nhwc), B, H * W,
                                              self._stream()), "gq_obs_to_state_train")
        self.launches += 1
        return (state, nhwc) if nhwc_bf16 else state

    def greedy(self, q):
        """flat arg-max per image and its split into (pixel index, rotation index) as transform_action does
        (Grasping_Agent_multidiscrete.py:295-299,381-386)"""
        t = self.torch
        B = q.shape[0]
        n = q[0].numel()
        idx = t.empty(B, dtype=t.int32, device=self.dev)
        val = t.empty(B, dtype=t.float32, device=self.dev)
        self._ck(self.L.gq_argmax(self._p(q.contiguous()), B, n, self._p(idx), self._p(val), self._stream()), "gq_argmax")
        hw = q.shape[2] * q.shape[3]
        return t.stack([idx % hw, idx // hw], dim=1), val
