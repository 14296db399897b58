"""Synthetic 12-bit tone + noise channels, generated on the GPU with torch.

x_c[n] = clip(round(A sin(2 pi f_c n / fs + phi_c) + sigma g_c[n]), -2048, 2047) as int16,
the range of the reference's sign-extended 12-bit XADC word (IMP/dsp_system_top.vhd:435);
f_c = 20 kHz + (c mod 4096) * 100 Hz, phi_c = 2 pi frac(c * 0.6180339887), fs = 1 MSPS."""
from __future__ import annotations

import math

FS_HZ = 1_000_000.0


def tone_noise(channels: int, n: int, device, seed: int = 0x5D12, first_channel: int = 0, frame: int = 0,
               amp: float = 1400.0, sigma: float = 100.0, block: int = 2048):
    """int16 [channels, n] on `device`; `frame` advances the time origin so that
    consecutive frames form one continuous stream."""
    import torch
    gen = torch.Generator(device=device)
    gen.manual_seed(seed + 7919 * frame)
    out = torch.empty((channels, n), dtype=torch.int16, device=device)
    t = (torch.arange(n, device=device, dtype=torch.float64) + float(frame) * n) / FS_HZ
    for c0 in range(0, channels, block):
        c1 = min(channels, c0 + block)
        c = torch.arange(first_channel + c0, first_channel + c1, device=device, dtype=torch.float64)
        f = 20e3 + torch.remainder(c, 4096.0) * 100.0
        phi = 2 * math.pi * torch.remainder(c * 0.6180339887, 1.0)
        ph = torch.remainder(f[:, None] * t[None, :], 1.0) * (2 * math.pi) + phi[:, None]
        sig = amp * torch.sin(ph).to(torch.float32)
        sig += sigma * torch.randn(sig.shape, generator=gen, device=device, dtype=torch.float32)
        out[c0:c1] = torch.clamp(torch.round(sig), -2048, 2047).to(torch.int16)
    return out


def full_range(channels: int, n: int, device, seed: int = 1):
    """Adversarial stimulus: uniform int16 including -32768 (bit-exactness tests)."""
    import torch
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    return torch.randint(-32768, 32768, (channels, n), generator=gen, device=device, dtype=torch.int32).to(torch.int16)
