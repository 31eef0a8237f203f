"""Deterministic parameter fill shared by tools/make_golden.py and the tests, so that golden fixtures
only need to carry (seed, checksum) instead of megabytes of weights.  TEST INFRASTRUCTURE ONLY."""
import torch


def fill_params(model, seed: int) -> float:
    """Overwrite every parameter (sorted by name) from a seeded CPU generator; returns a checksum."""
    g = torch.Generator().manual_seed(seed)
    sd = dict(model.named_parameters())
    chk = 0.0
    with torch.no_grad():
        for name in sorted(sd):
            p = sd[name]
            if p.dim() >= 2:
                fan_in = p.numel() // p.shape[0]
                val = (torch.rand(p.shape, generator=g) * 2 - 1) / (fan_in ** 0.5)
                if name.endswith("e_linear.weight"):
                    val = torch.randn(p.shape, generator=g) * 0.7 + 0.8
            elif name.endswith("skip"):
                val = torch.randn(p.shape, generator=g) * 0.7 + 0.3
            elif "norms" in name and name.endswith("weight"):
                val = 1.0 + 0.2 * torch.randn(p.shape, generator=g)
            else:
                val = 0.2 * torch.randn(p.shape, generator=g)
            if "relation_pri" in name:
                val = 1.0 + 0.5 * torch.randn(p.shape, generator=g)
            p.copy_(val.to(p.dtype))
            chk += float(val.double().abs().sum())
    return chk
