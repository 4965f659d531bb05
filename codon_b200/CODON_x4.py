"""Drop-in for the reference ``CODON_X4/CODON_x4.py``: exports ``CODONNet`` (see codon_b200/model.py)."""
from .model import CODONNetBase


class CODONNet(CODONNetBase):
    SCALE = 4
