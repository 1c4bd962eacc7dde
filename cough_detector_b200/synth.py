"""Seeded synthetic audio and weights (SURVEY.md §8d input recipe).

There is no network for datasets or checkpoints and the reference's shipped checkpoint was
trained on zeros (IMPROVEMENT_PLAN.md:19-24), so parity and benchmarks run on synthetic
16 kHz clips.  The mixture keeps the mel block from saturating at 1.0 (peak-normalised
broadband noise alone does) and follows the reference's own synthetic
generators: decaying noise bursts (setup_coughvid.py:383-405) and sine mixtures
(prepare_data.py:156-163).
"""
from __future__ import annotations

import math
from typing import Dict, List

import torch

SR = 16000


def _burst(g: torch.Generator, n: int) -> torch.Tensor:
    t = torch.arange(n, dtype=torch.float32) / SR
    onset = float(torch.rand((), generator=g)) * 0.6
    d = 0.3 + 0.5 * float(torch.rand((), generator=g))
    rel = (t - onset).clamp(min=0)
    env = torch.minimum(rel / 0.02, torch.ones(())) * torch.exp(-5.0 * rel / d) * (t >= onset)
    amp = 0.05 + 0.9 * float(torch.rand((), generator=g))
    return amp * env * torch.randn(n, generator=g) + 0.01 * torch.randn(n, generator=g)


def _sines(g: torch.Generator, n: int) -> torch.Tensor:
    t = torch.arange(n, dtype=torch.float32) / SR
    x = 0.005 * torch.randn(n, generator=g)
    for _ in range(1 + int(torch.randint(0, 3, (), generator=g))):
        f = 60.0 + 3940.0 * float(torch.rand((), generator=g))
        a = 0.01 + 0.49 * float(torch.rand((), generator=g))
        ph = 2 * math.pi * float(torch.rand((), generator=g))
        x = x + a * torch.sin(2 * math.pi * f * t + ph)
    return x


def synth_clips(n: int, seed: int = 0, fixed_length: bool = False) -> List[torch.Tensor]:
    """n clips as 1-D float32 CPU tensors: 25 % scaled noise, 25 % bursts, 25 % sines,
    15 % random-length (4000..40000) and 10 % edge cases (zero, impulse, DC, square, tiny, huge)."""
    g = torch.Generator().manual_seed(seed)
    out: List[torch.Tensor] = []
    for i in range(n):
        r = i % 20
        if r < 5:
            x = (0.001 + 0.299 * float(torch.rand((), generator=g))) * torch.randn(SR, generator=g)
        elif r < 10:
            x = _burst(g, SR)
        elif r < 15:
            x = _sines(g, SR)
        elif r < 18:
            ln = SR if fixed_length else int(torch.randint(4000, 40001, (), generator=g))
            x = _burst(g, ln) if r % 2 else _sines(g, ln)
        else:
            k = (i // 20) % 6 if r == 18 else ((i // 20) + 3) % 6
            if k == 0:
                x = torch.zeros(SR)
            elif k == 1:
                x = torch.zeros(SR)
                x[int(torch.randint(0, SR, (), generator=g))] = 0.7
            elif k == 2:
                x = torch.full((SR,), 0.25)
            elif k == 3:
                x = torch.sign(torch.sin(2 * math.pi * 440.0 * torch.arange(SR) / SR))
            elif k == 4:
                x = 1e-6 * torch.randn(SR, generator=g)
            else:
                x = 3.0e4 * _burst(g, SR)
        out.append(x.float().contiguous())
    return out


def edge_clips() -> List[torch.Tensor]:
    """The edge cases the domain has: silence, single impulse, DC, full-scale square, a very
    quiet and a very loud clip, a short (pad) and a long (centre-trim) clip, one sample."""
    g = torch.Generator().manual_seed(1234)
    sq = torch.sign(torch.sin(2 * math.pi * 440.0 * torch.arange(SR) / SR))
    imp = torch.zeros(SR)
    imp[777] = 0.7
    return [torch.zeros(SR), imp, torch.full((SR,), 0.25), sq,
            1e-6 * torch.randn(SR, generator=g), 3.0e4 * _burst(g, SR),
            _sines(g, 4001), _burst(g, 39999), torch.tensor([0.5]),
            torch.zeros(100), _sines(g, 16001), _burst(g, 15999)]


def synth_batch(b: int, seed: int = 0, device="cpu") -> torch.Tensor:
    """(b, 16000) fixed-length clips generated vectorised on ``device`` (bench inputs).
    Same four families as synth_clips, chosen by clip index; deterministic per (seed, device type)."""
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed)
    t = torch.arange(SR, device=dev, dtype=torch.float32) / SR
    out = torch.empty(b, SR, device=dev, dtype=torch.float32)
    step = 4096
    for s in range(0, b, step):
        m = min(step, b - s)
        u = torch.rand(m, 8, device=dev, generator=g)
        noise = torch.randn(m, SR, device=dev, generator=g)
        kind = (torch.arange(s, s + m, device=dev) % 4).view(-1, 1)
        flat = (0.001 + 0.299 * u[:, 0:1]) * noise
        rel = (t.view(1, -1) - 0.6 * u[:, 1:2]).clamp(min=0)
        env = torch.minimum(rel / 0.02, torch.ones((), device=dev)) * torch.exp(-5.0 * rel / (0.3 + 0.5 * u[:, 2:3]))
        env = env * (rel > 0)
        burst = (0.05 + 0.9 * u[:, 3:4]) * env * noise + 0.01 * noise.roll(1, dims=1)
        f1 = 60.0 + 3940.0 * u[:, 4:5]
        f2 = 60.0 + 3940.0 * u[:, 5:6]
        tone = (0.01 + 0.49 * u[:, 6:7]) * torch.sin(2 * math.pi * f1 * t) \
            + (0.01 + 0.49 * u[:, 7:8]) * torch.sin(2 * math.pi * f2 * t) + 0.005 * noise
        quiet = 0.02 * burst + 0.0005 * noise
        out[s:s + m] = torch.where(kind == 0, flat, torch.where(kind == 1, burst, torch.where(kind == 2, tone, quiet)))
    return out


def randomized_state_dict(model: torch.nn.Module, seed: int = 1, fc_scale: float = 20.0) -> Dict[str, torch.Tensor]:
    """state_dict with BatchNorm statistics / affine randomised and fc scaled so that BN folding
    is exercised and P(cough) spreads over (0, 1).  Applied on top of the
    model's current (default-init) weights; returns a new dict, the model is not modified."""
    g = torch.Generator().manual_seed(seed)
    sd = {k: v.detach().clone().cpu() for k, v in model.state_dict().items()}
    for k in list(sd):
        if k.endswith("running_mean"):
            p = k[: -len("running_mean")]
            c = sd[k].numel()
            sd[p + "weight"] = 0.5 + torch.rand(c, generator=g)
            sd[p + "bias"] = 0.1 * torch.randn(c, generator=g)
            sd[p + "running_mean"] = 0.1 * torch.randn(c, generator=g)
            sd[p + "running_var"] = 0.5 + torch.rand(c, generator=g)
    if "fc.2.weight" in sd:                      # CoughDetectorResidual's classifier
        sd["fc.2.weight"] = sd["fc.2.weight"] * fc_scale
    return sd


def spread_fc(sd: Dict[str, torch.Tensor], logits: torch.Tensor, centre: float = 0.8473,
              spread: float = 1.5, weight_key: str = "fc.2.weight", bias_key: str = "fc.2.bias") -> Dict[str, torch.Tensor]:
    """Rescale / recentre the classifier so P(cough) straddles 0.7 on the clips ``logits`` (computed
    with ``sd``) came from: logit margin d = l1 - l0 is mapped to spread*(d - median)/std + centre
    (centre = logit(0.7)).  Random-init GAP features barely move with the input, so without this
    every clip lands on the same side of the threshold and 'identical decisions' is vacuous."""
    sd = {k: v.clone() for k, v in sd.items()}
    b = sd[bias_key]
    d = (logits[:, 1] - logits[:, 0]) - (b[1] - b[0])
    s = spread / float(d.std().clamp(min=1e-6))
    sd[weight_key] = sd[weight_key] * s
    sd[bias_key] = torch.tensor([0.0, centre - s * float(d.median())])
    return sd
