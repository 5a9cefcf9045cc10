"""B200-native NGCF embedding-propagation hot path behind the reference's NGCF / BPR API."""
from .NGCF import NGCF
from .bprloss import BPR
from .graph import GraphedStep
from .optim import Adam
from .scoring import score_topk

__all__ = ["NGCF", "BPR", "GraphedStep", "Adam", "score_topk"]
