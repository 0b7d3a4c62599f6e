"""Parameter container with the reference UNet's exact state_dict surface.

Mirrors FastDiffSR/model/fastdiffsr_modules/unet.py:224-297 (module tree and key names, including
the 22 constructed-but-never-executed 1x1 `.conv` layers, unet.py:212) so that checkpoints written
by the reference load with strict=True.  It holds parameters only: the arithmetic of
UNet.forward (unet.py:299-323) is executed by libfdsr's CUDA kernels, never by torch modules.
"""
from __future__ import annotations

from torch import nn


class _Holder(nn.Module):
    """A module that only groups children / parameters."""


def _block(dim, dim_out, groups):
    b = _Holder()
    # indices 0 (GroupNorm) and 3 (Conv2d) carry parameters; 1 = Swish, 2 = Dropout in the reference
    b.block = nn.Sequential(nn.GroupNorm(groups, dim), nn.Identity(), nn.Identity(),
                            nn.Conv2d(dim, dim_out, 3, padding=1))
    return b


def _resnet_block(dim, dim_out, emb_dim, groups):
    r = _Holder()
    r.noise_func = _Holder()
    r.noise_func.noise_func = nn.Sequential(nn.Linear(emb_dim, dim_out))
    r.block1 = _block(dim, dim_out, groups)
    r.block2 = _block(dim_out, dim_out, groups)
    r.res_conv = nn.Conv2d(dim, dim_out, 1) if dim != dim_out else nn.Identity()
    return r


def _res_attn(dim, dim_out, emb_dim, groups, with_attn):
    m = _Holder()
    m.res_block = _resnet_block(dim, dim_out, emb_dim, groups)
    m.conv = nn.Conv2d(dim_out, dim_out, kernel_size=1, bias=True)  # dead in the reference too
    if with_attn:
        m.ca = _Holder()
        m.ca.fc1 = nn.Conv2d(dim_out, dim_out // 16, 1, bias=False)
        m.ca.fc2 = nn.Conv2d(dim_out // 16, dim_out, 1, bias=False)
        m.sa = _Holder()
        m.sa.conv1 = nn.Conv2d(2, 1, 7, padding=3, bias=False)
    return m


def _resample(dim, stride):
    m = _Holder()
    m.conv = nn.Conv2d(dim, dim, 3, stride, 1)
    return m


class UNet(nn.Module):
    def __init__(self, in_channel=6, out_channel=3, inner_channel=32, norm_groups=32, channel_mults=(1, 2, 4, 4),
                 attn_res=(8), res_blocks=3, dropout=0, with_noise_level_emb=True, image_size=256):
        super().__init__()
        self.cfg = dict(in_channel=in_channel, out_channel=out_channel, inner_channel=inner_channel,
                        norm_groups=norm_groups, channel_multiplier=list(channel_mults), attn_res=attn_res,
                        res_blocks=res_blocks, dropout=dropout)
        if not with_noise_level_emb:
            raise NotImplementedError("the sampling path always uses the noise-level embedding")
        emb = inner_channel
        self.noise_level_mlp = nn.Sequential(nn.Identity(), nn.Linear(inner_channel, inner_channel * 4),
                                             nn.Identity(), nn.Linear(inner_channel * 4, inner_channel))
        pre = inner_channel
        feat = [pre]
        downs = [nn.Conv2d(in_channel, inner_channel, kernel_size=3, padding=1)]
        n = len(channel_mults)
        for ind in range(n):
            cm = inner_channel * channel_mults[ind]
            for _ in range(res_blocks):
                downs.append(_res_attn(pre, cm, emb, norm_groups, False))  # attn_res is ignored (unet.py:261)
                feat.append(cm)
                pre = cm
            if ind != n - 1:
                downs.append(_resample(pre, 2))
                feat.append(pre)
        self.downs = nn.ModuleList(downs)
        self.mid = nn.ModuleList([_res_attn(pre, pre, emb, norm_groups, True),
                                  _res_attn(pre, pre, emb, norm_groups, False)])
        ups = []
        for ind in reversed(range(n)):
            cm = inner_channel * channel_mults[ind]
            for _ in range(res_blocks + 1):
                ups.append(_res_attn(pre + feat.pop(), cm, emb, norm_groups, False))
                pre = cm
            if ind >= 1:
                ups.append(_resample(pre, 1))
        self.ups = nn.ModuleList(ups)
        self.final_conv = _block(pre, out_channel if out_channel is not None else in_channel, norm_groups)

    def forward(self, x, time):
        raise RuntimeError("UNet.forward runs inside libfdsr; call GaussianDiffusion.denoise(...) or "
                           "super_resolution(...) on the owning netG")


# ------------------------------------------------------------------------------------------------------------
# SR3 baseline (which_model_G == "ddpm"): parameter container with the state_dict surface of
# FastDiffSR/model/ddpm_modules/unet.py:150-225 (time_mlp with its inv_freq buffer, ResnetBlock.mlp,
# SelfAttention norm / qkv / out).  Arithmetic again runs in libfdsr.
# ------------------------------------------------------------------------------------------------------------

class _TimeEmbedding(nn.Module):
    def __init__(self, dim):
        super().__init__()
        import math
        import torch
        self.dim = dim
        self.register_buffer("inv_freq", torch.exp(torch.arange(0, dim, 2, dtype=torch.float32) *
                                                   (-math.log(10000) / dim)))


def _sr3_resnet_block(dim, dim_out, emb_dim, groups):
    r = _Holder()
    r.mlp = nn.Sequential(nn.Identity(), nn.Linear(emb_dim, dim_out))  # index 0 = Swish in the reference
    r.block1 = _block(dim, dim_out, groups)
    r.block2 = _block(dim_out, dim_out, groups)
    r.res_conv = nn.Conv2d(dim, dim_out, 1) if dim != dim_out else nn.Identity()
    return r


def _sr3_res_attn(dim, dim_out, emb_dim, groups, with_attn):
    m = _Holder()
    m.res_block = _sr3_resnet_block(dim, dim_out, emb_dim, groups)
    if with_attn:
        m.attn = _Holder()
        m.attn.norm = nn.GroupNorm(groups, dim_out)
        m.attn.qkv = nn.Conv2d(dim_out, dim_out * 3, 1, bias=False)
        m.attn.out = nn.Conv2d(dim_out, dim_out, 1)
    return m


class SR3UNet(nn.Module):
    def __init__(self, in_channel=6, out_channel=3, inner_channel=32, norm_groups=32, channel_mults=(1, 2, 4, 8, 8),
                 attn_res=(8,), res_blocks=3, dropout=0, with_time_emb=True, image_size=128):
        super().__init__()
        attn_res = list(attn_res)  # (`now_res in attn_res` in the reference: a JSON list)
        self.cfg = dict(in_channel=in_channel, out_channel=out_channel, inner_channel=inner_channel,
                        norm_groups=norm_groups, channel_multiplier=list(channel_mults), attn_res=attn_res,
                        res_blocks=res_blocks, dropout=dropout, model="ddpm", image_size=image_size)
        if not with_time_emb:
            raise NotImplementedError("the sampling path always uses the time embedding")
        emb = inner_channel
        self.time_mlp = nn.Sequential(_TimeEmbedding(inner_channel), nn.Linear(inner_channel, inner_channel * 4),
                                      nn.Identity(), nn.Linear(inner_channel * 4, inner_channel))
        pre = inner_channel
        feat = [pre]
        now = image_size
        downs = [nn.Conv2d(in_channel, inner_channel, kernel_size=3, padding=1)]
        n = len(channel_mults)
        for ind in range(n):
            cm = inner_channel * channel_mults[ind]
            for _ in range(res_blocks):
                downs.append(_sr3_res_attn(pre, cm, emb, norm_groups, now in attn_res))
                feat.append(cm)
                pre = cm
            if ind != n - 1:
                downs.append(_resample(pre, 2))
                feat.append(pre)
                now //= 2
        self.downs = nn.ModuleList(downs)
        self.mid = nn.ModuleList([_sr3_res_attn(pre, pre, emb, norm_groups, True),
                                  _sr3_res_attn(pre, pre, emb, norm_groups, False)])
        ups = []
        for ind in reversed(range(n)):
            cm = inner_channel * channel_mults[ind]
            for _ in range(res_blocks + 1):
                ups.append(_sr3_res_attn(pre + feat.pop(), cm, emb, norm_groups, now in attn_res))
                pre = cm
            if ind >= 1:
                ups.append(_resample(pre, 1))
                now *= 2
        self.ups = nn.ModuleList(ups)
        self.final_conv = _block(pre, out_channel if out_channel is not None else in_channel, norm_groups)

    def forward(self, x, time):
        raise RuntimeError("UNet.forward runs inside libfdsr; call GaussianDiffusion.denoise(...) or "
                           "super_resolution(...) on the owning netG")
