"""Drop-in for the reference ``CODON_X16/CODON_x16.py``: exports ``CODONNet`` (see codon_b200/model.py)."""
from .model import CODONNetBase


class CODONNet(CODONNetBase):
    SCALE = 16
