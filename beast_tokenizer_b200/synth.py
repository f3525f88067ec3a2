"""Seeded synthetic trajectories: per (trajectory, DoF) a
sinusoid with amplitude 0.02*N(0,1), frequency U(0,3), phase U(0,2pi), plus
0.002*N(0,1) noise, sampled at t = linspace(0, 1, T).  Used by the tests, the
golden-fixture generator and bench.py; there is no dataset in the image."""
import math

import torch


def synth(batch: int, seq_len: int, num_dof: int, seed: int, device="cpu") -> torch.Tensor:
    """fp32 [batch, seq_len, num_dof].  The random draws are made on the CPU
    generator so the same seed gives the same trajectories on every box; the
    elementwise arithmetic runs on `device`."""
    g = torch.Generator(device="cpu").manual_seed(int(seed))
    amp = 0.02 * torch.randn(batch, 1, num_dof, generator=g)
    freq = 3.0 * torch.rand(batch, 1, num_dof, generator=g)
    phase = 2.0 * math.pi * torch.rand(batch, 1, num_dof, generator=g)
    noise = 0.002 * torch.randn(batch, seq_len, num_dof, generator=g)
    t = torch.linspace(0.0, 1.0, seq_len).view(1, seq_len, 1)
    amp, freq, phase, noise, t = (x.to(device) for x in (amp, freq, phase, noise, t))
    return (amp * torch.sin(2.0 * math.pi * freq * t + phase) + noise).to(torch.float32).contiguous()


class SyntheticLoader:
    """Iterable of `{"actions": [batch, seq_len, num_dof]}` dicts, the shape the
    reference's `fit_parameters` / `fit_from_trajectories` consume
    (beast/beast_bspline_tokenizer.py:193-197, beast/beast_bpe_trainer.py:122-129)."""

    def __init__(self, num_batches, batch, seq_len, num_dof, seed0=0, device="cpu", key="actions"):
        self.num_batches, self.batch, self.seq_len, self.num_dof = num_batches, batch, seq_len, num_dof
        self.seed0, self.device, self.key = seed0, device, key

    def __len__(self):
        return self.num_batches

    def __iter__(self):
        for i in range(self.num_batches):
            yield {self.key: synth(self.batch, self.seq_len, self.num_dof, self.seed0 + i, self.device)}


def synth_device(batch: int, seq_len: int, num_dof: int, seed: int, device) -> torch.Tensor:
    """Same distribution as `synth`, every draw made by the device generator (bench-scale inputs:
    1.6 M trajectories would take minutes from the CPU generator)."""
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(int(seed))
    amp = 0.02 * torch.randn(batch, 1, num_dof, generator=g, device=dev)
    freq = 3.0 * torch.rand(batch, 1, num_dof, generator=g, device=dev)
    phase = 2.0 * math.pi * torch.rand(batch, 1, num_dof, generator=g, device=dev)
    noise = 0.002 * torch.randn(batch, seq_len, num_dof, generator=g, device=dev)
    t = torch.linspace(0.0, 1.0, seq_len, device=dev).view(1, seq_len, 1)
    return (amp * torch.sin(2.0 * math.pi * freq * t + phase) + noise).to(torch.float32).contiguous()
