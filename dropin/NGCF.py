"""Drop-in for the reference's ``model/NGCF.py``: put this directory on ``sys.path`` ahead of ``model/`` (or copy this
file over ``model/NGCF.py``) and ``from NGCF import NGCF`` (main.py:9, demo.py:5) resolves to the B200 module."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.append(_ROOT)

from seoul_tourism_recommendation_ngcf_b200.NGCF import NGCF  # noqa: E402,F401
