"""Drop-in replacements of the reference's models.HEATNet2 / HEATNet4 / HGT (same constructor and
forward(G, h=None) signatures, same state_dict keys and shapes)."""
from .heat import HEATLayer, HEATNet2, HEATNet4
from .hgt import HGT, HGTLayer

__all__ = ["HEATLayer", "HEATNet2", "HEATNet4", "HGT", "HGTLayer"]
