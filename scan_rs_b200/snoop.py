"""Progress / cancel tokens mirroring the reference's `snoop` crate (snoop/src/lib.rs)."""
from __future__ import annotations

import threading


class NoOpSnoop:
    """snoop::NoOpSnoop (snoop/src/lib.rs:60-85): never cancelled, ignores progress."""

    def is_cancelled(self) -> bool:
        return False

    def set_progress(self, fraction: float) -> None:
        pass


class AtomicSnoop:
    """snoop::AtomicSnoop (snoop/src/lib.rs:87-226), without sub-snoop scaling: a shared cancel
    flag another thread can set and a progress fraction it can read."""

    def __init__(self):
        self._cancel = threading.Event()
        self._progress = 0.0
        self.history = []

    def cancel(self) -> None:
        self._cancel.set()

    def is_cancelled(self) -> bool:
        return self._cancel.is_set()

    def set_progress(self, fraction: float) -> None:
        self._progress = fraction
        self.history.append(fraction)

    def get_progress(self) -> float:
        return self._progress
