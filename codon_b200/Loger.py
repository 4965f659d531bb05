"""Console + file tee with the public surface of the reference's ``Loger.Logger`` (CODON_X4/Loger.py:22-57):
``sys.stdout = Logger(path)`` as in CODON_X4/test.py:53, then ``print`` goes to both.

Own implementation: a list of sinks instead of two hard-wired streams, ``pathlib`` for the directory, durable flush
(``fsync``) only for real files, idempotent ``close`` that never closes the interpreter's console (the reference's
``close`` does, Loger.py:55, which breaks every later ``print``), and no matplotlib / torchvision imports.
"""
from __future__ import annotations

import os
import sys
from pathlib import Path
from typing import IO, List, Optional


def mkdir_if_missing(directory) -> None:
    """Creates ``directory`` (and parents); an empty name or an existing directory is fine."""
    if directory:
        Path(directory).mkdir(parents=True, exist_ok=True)


class Logger:
    """File-like object that duplicates everything written to it."""

    def __init__(self, fpath: Optional[str] = None, mode: str = "a"):
        self.fpath = fpath
        self.console: IO[str] = sys.stdout          # whatever stdout is NOW (so loggers can be stacked)
        self.file: Optional[IO[str]] = None
        if fpath is not None:
            mkdir_if_missing(Path(fpath).parent)
            self.file = open(fpath, mode)

    # -- sinks ---------------------------------------------------------------------------------------
    def _sinks(self) -> List[IO[str]]:
        return [s for s in (self.console, self.file) if s is not None]

    # -- file protocol -------------------------------------------------------------------------------
    def write(self, msg: str) -> int:
        for sink in self._sinks():
            sink.write(msg)
        return len(msg)

    def writelines(self, lines) -> None:
        for line in lines:
            self.write(line)

    def flush(self) -> None:
        for sink in self._sinks():
            sink.flush()
        if self.file is not None:
            os.fsync(self.file.fileno())             # the reference fsyncs the log on every flush as well

    def isatty(self) -> bool:
        return False

    def close(self) -> None:
        f, self.file = self.file, None
        if f is not None and not f.closed:
            f.close()

    # -- context manager / finaliser -----------------------------------------------------------------
    def __enter__(self) -> "Logger":
        return self

    def __exit__(self, *exc) -> None:
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
