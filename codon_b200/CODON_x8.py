"""Drop-in for the reference ``CODON_X8/CODON_x8.py``: exports ``CODONNet`` (see codon_b200/model.py)."""
from .model import CODONNetBase


class CODONNet(CODONNetBase):
    SCALE = 8
