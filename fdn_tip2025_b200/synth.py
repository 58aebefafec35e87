"""Seeded synthetic weights and synthetic low-light inputs.

The FDN checkpoints are absent from the reference mount, so parity and timing
use synthetic state_dicts that follow the exact key/shape schema (schema.py).  The generator only
depends on ``torch.Generator`` on CPU, which is bit-reproducible across machines, so the GPU box
regenerates the very same tensors the golden fixtures were made with.

``damp`` multiplies every ``net_p.*project_out.weight``: with 0.03 the network is
well conditioned (reference fp32-vs-fp64 3e-6) while every code path is still exercised.
"""
import math

import torch

from . import schema


def _fill(shape, kind, gen, dtype):
    if kind == "ones":
        # LayerNorm / spectral gains: perturb around 1 so a wrong gain index cannot hide
        return 1.0 + 0.2 * (torch.rand(shape, generator=gen, dtype=torch.float64) - 0.5)
    if kind == "zeros":
        return 0.2 * (torch.rand(shape, generator=gen, dtype=torch.float64) - 0.5)
    if kind == "conv":
        fan_in = shape[1] * shape[2] * shape[3]
        bound = 1.0 / math.sqrt(fan_in)
        return (torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * bound
    if kind == "linear":
        bound = 1.0 / math.sqrt(shape[1])
        return (torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * bound
    if kind.startswith("conv_bias:"):
        bound = 1.0 / math.sqrt(int(kind.split(":")[1]))
        return (torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * bound
    if kind == "buf_zeros":     # BN running_mean
        return 0.1 * (torch.rand(shape, generator=gen, dtype=torch.float64) - 0.5)
    if kind == "buf_ones":      # BN running_var
        return 0.5 + torch.rand(shape, generator=gen, dtype=torch.float64)
    raise ValueError(kind)


def make_state_dict(table, seed=0, damp=None, dtype=torch.float32):
    """table: OrderedDict key -> (shape, kind) from schema.py."""
    gen = torch.Generator().manual_seed(seed)
    sd = {}
    for key, (shape, kind) in table.items():
        if kind == "buf_long":
            sd[key] = torch.tensor(1000, dtype=torch.int64)
            continue
        t = _fill(shape, kind, gen, dtype)
        if damp is not None and key.startswith("net_p.") and key.endswith("project_out.weight"):
            t = t * damp
        sd[key] = t.to(dtype).contiguous()
    return sd


def fdn_state_dict(dim=32, seed=0, damp=0.03, dtype=torch.float32):
    return make_state_dict(schema.fdn_schema(dim), seed=seed, damp=damp, dtype=dtype)


def mar_state_dict(seed=0, dtype=torch.float32):
    return make_state_dict(schema.mar_schema(), seed=seed, dtype=dtype)


def lpnet_state_dict(seed=0, dtype=torch.float32):
    return make_state_dict(schema.lpnet_schema(), seed=seed, dtype=dtype)


def low_light_images(b, h, w, first_index=0, dtype=torch.float32):
    """Synthetic low-light frames: smooth base + sensor noise, in [0, 0.22]."""
    imgs = []
    for i in range(b):
        gen = torch.Generator().manual_seed(1000 + first_index + i)
        base = torch.rand(1, 3, max(h // 16, 2), max(w // 16, 2), generator=gen)
        base = torch.nn.functional.interpolate(base, size=(h, w), mode="bicubic", align_corners=False)
        noise = torch.rand(1, 3, h, w, generator=gen)
        imgs.append((0.2 * base + 0.02 * noise).clamp_(0.0, 1.0))
    return torch.cat(imgs, 0).to(dtype).contiguous()
