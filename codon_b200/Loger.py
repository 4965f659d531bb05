"""Drop-in for the reference ``Loger`` (CODON_X4/Loger.py:22-57): ``Logger(fpath)`` tees writes to
the console and to a file, flushing both (with fsync) on ``flush``.  Unlike the reference it does not
import matplotlib / torchvision / torch.distributed, and ``close`` leaves the real console open
(the reference closes it, Loger.py:55)."""
import os
import sys


def mkdir_if_missing(dir_path):
    if dir_path:
        os.makedirs(dir_path, exist_ok=True)


class Logger(object):
    def __init__(self, fpath=None):
        self.console = sys.stdout
        self.file = None
        self.fpath = fpath
        if fpath is not None:
            mkdir_if_missing(os.path.dirname(fpath))
            self.file = open(fpath, "a")

    def __del__(self):
        self.close()

    def __enter__(self):
        return self

    def __exit__(self, *args):
        self.close()

    def write(self, msg):
        self.console.write(msg)
        if self.file is not None:
            self.file.write(msg)

    def flush(self):
        self.console.flush()
        if self.file is not None:
            self.file.flush()
            os.fsync(self.file.fileno())

    def close(self):
        if self.file is not None:
            self.file.close()
            self.file = None
