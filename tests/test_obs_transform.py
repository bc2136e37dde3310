"""Training-time observation transform on the device: Grasp_Agent.transform_observation(normalize=True, jitter_and_noise=True)
— Grasping_Agent_multidiscrete.py:301-368 with the transforms of :118-124 — batched in libgrasp_qnet.so (gq_obs_to_state_train).

Oracle: torchvision's own functional colour operations (the arithmetic ColorJitter applies to float tensors) with the same factors
and operation order, and a numpy restatement of the counter-based noise generator (Philox4x32-10 + Box-Muller).  The reference runs
ColorJitter on a PIL image, which quantises to 8 bits after every operation; the device works in float like torchvision's tensor
path, so "same as the reference" here means within the 8-bit quantisation steps of the PIL path (checked as such below).
"""
import numpy as np
import pytest
import torch

gpu = pytest.mark.gpu


# ------------------------------------------------------------------------------------------------ numpy mirrors of the kernel arithmetic
def mirror_jitter(img, fac, order):
    """img [H,W,3] float32 in [0,1]; same operation sequence as jitter_op / k_obs_train_state"""
    f32 = np.float32
    x = img.astype(f32).copy()
    gray = lambda a: (f32(0.2989) * a[..., 0] + f32(0.587) * a[..., 1] + f32(0.114) * a[..., 2]).astype(f32)
    for op in order:
        f = f32(fac[op])
        if op == 0:
            x = np.clip(f * x, 0, 1)
        elif op == 1:
            x = np.clip(f * x + (f32(1) - f) * f32(gray(x).astype(np.float64).mean()), 0, 1)
        elif op == 2:
            x = np.clip(f * x + ((f32(1) - f) * gray(x))[..., None], 0, 1)
        else:
            r, g, b = x[..., 0], x[..., 1], x[..., 2]
            maxc, minc = x.max(-1), x.min(-1)
            eqc = maxc == minc
            cr = maxc - minc
            s = cr / np.where(eqc, f32(1), maxc)
            div = np.where(eqc, f32(1), cr)
            rc, gc, bc = (maxc - r) / div, (maxc - g) / div, (maxc - b) / div
            hr = (maxc == r) * (bc - gc)
            hg = ((maxc == g) & (maxc != r)) * (f32(2) + rc - bc)
            hb = ((maxc != g) & (maxc != r)) * (f32(4) + gc - rc)
            h = np.fmod((hr + hg + hb) / f32(6) + f32(1), f32(1)).astype(f32)
            h = h + f
            h = (h - np.floor(h)).astype(f32)
            v = maxc
            h6 = h * f32(6)
            fl = np.floor(h6)
            ff = (h6 - fl).astype(f32)
            i = fl.astype(np.int32) % 6
            p = np.clip(v * (1 - s), 0, 1)
            q = np.clip(v * (1 - s * ff), 0, 1)
            t = np.clip(v * (1 - s * (1 - ff)), 0, 1)
            sel = lambda a: np.choose(i, a).astype(f32)
            x = np.stack([sel([v, q, p, p, t, v]), sel([t, v, v, q, p, p]), sel([p, p, t, v, v, q])], axis=-1)
        x = x.astype(f32)
    return x


def torchvision_jitter(img, fac, order):
    import torchvision.transforms.functional as F

    x = torch.from_numpy(img.astype(np.float32)).permute(2, 0, 1)
    for op in order:
        f = float(fac[op])
        x = [F.adjust_brightness, F.adjust_contrast, F.adjust_saturation, F.adjust_hue][op](x, f)
    return x.permute(1, 2, 0).numpy()


def philox_normal(seed, env, step, pixels):
    """numpy restatement of obs_noise(): Philox4x32-10, counter (pixel, env_lo, env_hi, step), key (seed_lo, seed_hi)"""
    M0, M1, W0, W1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
    c0 = pixels.astype(np.uint32)
    c1 = np.full_like(c0, np.uint32(env & 0xFFFFFFFF))
    c2 = np.full_like(c0, np.uint32((env >> 32) & 0xFFFFFFFF))
    c3 = np.full_like(c0, np.uint32(step))
    k0, k1 = np.uint32(seed & 0xFFFFFFFF), np.uint32((seed >> 32) & 0xFFFFFFFF)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0, k1 = np.uint32(k0 + W0), np.uint32(k1 + W1)
    u1 = ((c0 >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)
    u2 = ((c1 >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)
    return np.sqrt(-2.0 * np.log(u1.astype(np.float64))) * np.cos(2.0 * np.pi * u2.astype(np.float64))


# ------------------------------------------------------------------------------------------------ CPU
def test_mirror_of_the_kernel_arithmetic_equals_torchvision():
    """every permutation of the four operations, random factors in ColorJitter(0.5, 0.5, 0.5, 0.5)'s ranges"""
    import itertools

    rng = np.random.RandomState(0)
    img = rng.randint(0, 256, (40, 40, 3)).astype(np.float32) / 255.0
    img[:4] = 0.5  # grey pixels: maxc == minc branch of the hue conversion
    worst = 0.0
    for order in itertools.permutations(range(4)):
        fac = [rng.uniform(0.5, 1.5), rng.uniform(0.5, 1.5), rng.uniform(0.5, 1.5), rng.uniform(-0.5, 0.5)]
        worst = max(worst, float(np.abs(mirror_jitter(img, fac, order) - torchvision_jitter(img, fac, order)).max()))
    assert worst < 5e-6, worst


def test_float_jitter_is_within_quantisation_of_the_reference_pil_path():
    """the reference applies the same operations to a PIL image (8-bit after every operation, Grasping_Agent_multidiscrete.py:118-124):
    the float path stays within a few grey levels of it"""
    import torchvision.transforms.functional as F
    from PIL import Image

    rng = np.random.RandomState(1)
    u8 = rng.randint(0, 256, (32, 32, 3)).astype(np.uint8)
    fac, order = [1.3, 0.7, 1.2, 0.1], (2, 0, 3, 1)
    pil = Image.fromarray(u8)
    for op in order:
        pil = [F.adjust_brightness, F.adjust_contrast, F.adjust_saturation, F.adjust_hue][op](pil, fac[op])
    ours = mirror_jitter(u8.astype(np.float32) / 255.0, fac, order) * 255.0
    diff = np.abs(ours - np.asarray(pil, dtype=np.float32))
    assert np.median(diff) <= 1.5 and np.percentile(diff, 99) <= 8.0, (np.median(diff), np.percentile(diff, 99))


def test_noise_generator_mirror_is_standard_normal_and_counter_based():
    z = philox_normal(1234, 7, 3, np.arange(40000))
    assert abs(z.mean()) < 0.02 and abs(z.std() - 1.0) < 0.02 and abs(((z - z.mean()) ** 3).mean()) < 0.05
    assert np.array_equal(z[100:200], philox_normal(1234, 7, 3, np.arange(100, 200)))          # value depends on the counter only
    assert not np.allclose(z[:100], philox_normal(1234, 8, 3, np.arange(100)))                 # env, step and seed all matter
    assert not np.allclose(z[:100], philox_normal(1234, 7, 4, np.arange(100)))
    assert not np.allclose(z[:100], philox_normal(1235, 7, 3, np.arange(100)))
    assert abs(np.corrcoef(z[:-1], z[1:])[0, 1]) < 0.02


# ------------------------------------------------------------------------------------------------ GPU
def _obs(B, seed=0):
    g = torch.Generator().manual_seed(seed)
    rgb = torch.randint(0, 256, (B, 200, 200, 3), generator=g, dtype=torch.uint8)
    depth = 0.9 + 0.3 * torch.rand((B, 200, 200), generator=g)
    return {"rgb": rgb.cuda(), "depth": depth.cuda()}


@gpu
def test_training_transform_matches_torchvision_and_the_noise_mirror():
    from mujoco_rl_ur5_b200.qnet import QNetForward, make_torch_qnet

    torch.manual_seed(0)
    qf = QNetForward(make_torch_qnet(6).state_dict(), 0, max_batch=8)
    B = 6
    obs = _obs(B)
    torch.manual_seed(5)
    fac, order = qf.draw_color_jitter(B)
    env_index = torch.tensor([3, 9, 4096, 17, 2 ** 33 + 5, 0])
    state, nhwc = qf.obs_to_state_train(obs, 1.1, 0.001, seed=77, step=12, env_index=env_index, jitter=fac, order=order, nhwc_bf16=True)
    state = state.cpu().numpy()
    rgb, depth = obs["rgb"].cpu().numpy(), obs["depth"].cpu().numpy()
    for b in range(B):
        want = torchvision_jitter(rgb[b].astype(np.float32) / 255.0, fac[b].cpu().numpy(), order[b].cpu().numpy().tolist())
        got = state[b, :3].transpose(1, 2, 0)
        assert np.abs(got - want).max() < 2e-5, (b, np.abs(got - want).max())
        z = philox_normal(77, int(env_index[b]), 12, np.arange(40000)).reshape(200, 200)
        d = -(np.minimum(depth[b], np.float32(1.1)).astype(np.float64) + 0.001 * z)
        d = (d - d.min()) / (d.max() - d.min())
        assert np.abs(state[b, 3] - d).max() < 2e-5, (b, np.abs(state[b, 3] - d).max())
    # the NHWC bf16 output is the same tensor, rounded
    assert torch.equal(nhwc.cpu(), torch.from_numpy(state).permute(0, 2, 3, 1).to(torch.bfloat16))
    # a pixel's value does not depend on how the envs are batched: rows 2..5 alone, with their global env ids
    sub = {"rgb": obs["rgb"][2:], "depth": obs["depth"][2:]}
    part = qf.obs_to_state_train(sub, 1.1, 0.001, seed=77, step=12, env_index=env_index[2:], jitter=fac[2:].contiguous(), order=order[2:].contiguous())
    assert np.array_equal(part.cpu().numpy(), state[2:])
    # without jitter and noise it is the deterministic transform the acting agent uses
    plain = qf.obs_to_state_train(obs, 1.1, 0.0, seed=1, step=0)
    assert torch.allclose(plain, qf.obs_to_state(obs, 1.1), atol=1e-6)


@gpu
def test_depth_noise_statistics_on_device():
    from mujoco_rl_ur5_b200.qnet import QNetForward, make_torch_qnet

    torch.manual_seed(0)
    qf = QNetForward(make_torch_qnet(6).state_dict(), 0, max_batch=8)
    obs = {"rgb": torch.zeros((4, 200, 200, 3), dtype=torch.uint8, device="cuda"),
           "depth": torch.linspace(0.9, 1.0, 40000, device="cuda").reshape(1, 200, 200).repeat(4, 1, 1).contiguous()}
    a = qf.obs_to_state_train(obs, 1.1, 0.001, seed=3, step=1)[:, 3]
    # the clean input is a linear ramp spanning 0.1 m, so the normalised output is a line + noise / (range of the noisy image ~ 0.106 m)
    y = a[0].reshape(-1).double().cpu().numpy()
    x = np.arange(y.size)
    res = y - np.polyval(np.polyfit(x, y, 1), x)
    assert 0.0085 < res.std() < 0.0105, res.std()
    assert abs(np.corrcoef(res[:-1], res[1:])[0, 1]) < 0.03
    c = qf.obs_to_state_train(obs, 1.1, 0.001, seed=3, step=2)[:, 3]
    assert not torch.allclose(a, c)                      # a new step draws new noise
    assert not torch.allclose(a[0], a[1])                # and every env its own
