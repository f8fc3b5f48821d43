"""Minimal logger shim and run-length helper used by the prediction path.

Mirrors the call surface of the reference's ``Messenger`` (``src/orcAI/auxiliary.py:29-200``)
because every hot-path function takes a ``msgr``; the cosmetic parts (platform / TensorFlow
device / memory printers) are out of scope (component 7).
``find_consecutive_ones`` mirrors ``auxiliary.py:420-440`` on top of the CUDA scan kernel.
"""

from __future__ import annotations

import time
from datetime import datetime, timedelta
from pathlib import Path

import click
import numpy as np

from orcai_b200 import __version__


class Messenger:
    """Verbosity-gated, indented console logger (0 errors, 1 warnings, 2 info, 3 debug)."""

    def __init__(
        self,
        title: str | None = None,
        n_indent: int = 0,
        verbosity: int = 2,
        indent_str: str = "    ",
        show_part_times: bool = True,
        file: Path | None = None,
    ):
        self.n_indent = n_indent
        self.verbosity = verbosity
        self.file = file
        self.indent_str = indent_str
        self.show_part_times = show_part_times
        self.start_time = time.time()
        self.part_times: list[float] = []
        if title is not None:
            self.start(title, severity=2)

    def print(self, message, indent: int = 0, set_indent: int | None = None, prepend: str = "", severity: int = 2, **style):
        if self.verbosity < severity:
            return
        if set_indent is not None:
            self.n_indent = set_indent
        if isinstance(message, dict):
            text = "\n".join(f"{self.indent_str * self.n_indent}{k}: {v}" for k, v in message.items())
        elif isinstance(message, (list, tuple)):
            text = "\n".join(f"{self.indent_str * self.n_indent}{v}" for v in message)
        else:
            text = self.indent_str * self.n_indent + prepend + str(message)
        click.echo(click.style(text, **style), file=self.file)
        self.n_indent += indent

    def debug(self, message, indent=0, set_indent=None, severity=3, **style):
        self.print(message, indent, set_indent, severity=severity, **style)

    def info(self, message, indent=0, set_indent=None, severity=2, **style):
        self.print(message, indent, set_indent, severity=severity, **style)

    def start(self, message, indent=0, set_indent=0, severity=2, **style):
        self.print(message, indent, set_indent, prepend="🐳 ", severity=severity, bold=True, **style)
        stamp = datetime.now().strftime("%Y-%m-%d %H:%M:%S")
        self.print(f"orcai_b200 {__version__} [started @ {stamp}]", indent, set_indent, severity=severity, italic=True, **style)

    def part(self, message, indent=1, set_indent=0, severity=2, **style):
        now = time.time()
        last = self.part_times.pop() if self.part_times else None
        self.part_times.append(now)
        if self.show_part_times:
            total = timedelta(seconds=round(now - self.start_time))
            delta = f", 𝚫 {timedelta(seconds=round(now - last))}" if last else ""
            message = f"{message} [{total}{delta}]"
        self.print(message, indent, set_indent, prepend="🐳 ", severity=severity, bold=True, **style)

    def success(self, message, indent=0, set_indent=0, severity=2, **style):
        self.part(message, indent, set_indent, severity=severity, fg="green", **style)

    def warning(self, message, indent=0, set_indent=None, severity=1, **style):
        self.print(message, indent, set_indent, prepend="‼️ ", severity=severity, fg="yellow", **style)

    def error(self, message, indent=0, set_indent=None, severity=0, **style):
        self.print(message, indent, set_indent, prepend="❌ ", severity=severity, fg="red", **style)


def find_consecutive_ones(binary_vector: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """Start and inclusive stop indices of the runs of ones in a 0/1 vector (GPU scan kernel)."""
    from orcai_b200.runtime import default_context

    v = np.asarray(binary_vector)
    ctx = default_context()
    # the segment kernel thresholds agg > threshold / max(count): feed the vector as "aggregate" with count 1
    _, starts, stops = ctx.threshold_segments(v.astype(np.float64).reshape(-1, 1), np.ones(len(v)), threshold=0.5)
    return starts.astype(np.int64), stops.astype(np.int64)
