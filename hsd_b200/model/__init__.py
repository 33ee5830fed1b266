from .HSD import HSD
from .multiscale_HSD import MultiHSD
from .dynamic_HSD import DynamicHSD
from .GraphWave import GraphWave

name = "model"
__all__ = ["HSD", "MultiHSD", "DynamicHSD", "GraphWave"]
