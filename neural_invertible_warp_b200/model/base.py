"""reference model/base.py:191-211 -- the Graph base class (losses only; the engine ``Model`` is
host orchestration and stays the reference's own)."""
from ._core import RenderCore, edict  # noqa: F401


class Graph(RenderCore):
    def __init__(self, opt, tb=None):
        super().__init__()

    def forward(self, opt, var, mode=None):
        raise NotImplementedError

    def compute_loss(self, opt, var, mode=None):
        raise NotImplementedError
