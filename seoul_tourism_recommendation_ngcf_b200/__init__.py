"""B200-native NGCF embedding-propagation hot path behind the reference's NGCF / BPR API."""
from .NGCF import NGCF
from .bprloss import BPR
from .scoring import score_topk

__all__ = ["NGCF", "BPR", "score_topk"]
