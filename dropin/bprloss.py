"""Drop-in for the reference's ``model/bprloss.py``: ``from bprloss import BPR`` (main.py:10) resolves to the B200
fused gather-dot-logsigmoid loss."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.append(_ROOT)

from seoul_tourism_recommendation_ngcf_b200.bprloss import BPR  # noqa: E402,F401
