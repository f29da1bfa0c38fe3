"""Go's text/tabwriter, as cmd/inspect uses it: NewWriter(os.Stdout, minwidth 0, tabwidth 0, padding 2,
padchar ' ', flags 0) (cmd/inspect/display.go:16,142). Elastic tabstops: a cell is text terminated by a
tab; the text after a line's last tab is a trailing cell that takes part in no column; a column block is
a run of consecutive lines that all have a tab-terminated cell in that column; its width is the widest
cell + padding; cells are left-aligned and padded with spaces."""
from __future__ import annotations

from typing import List


class Writer:
    def __init__(self, minwidth: int = 0, tabwidth: int = 0, padding: int = 2, padchar: str = " ", flags: int = 0):
        assert flags == 0 and padchar != "\t", "only the configuration cmd/inspect uses is restated"
        self.minwidth, self.padding, self.padchar = minwidth, padding, padchar
        self._text: List[str] = []

    def write(self, s: str) -> None:
        self._text.append(s)

    def flush(self) -> str:
        text = "".join(self._text)
        self._text = []
        ends_with_newline = text.endswith("\n")
        raw = text.split("\n")
        if ends_with_newline:
            raw.pop()
        lines = [ln.split("\t") for ln in raw]  # cells; the last one is the trailing (non tab-terminated) cell
        out: List[str] = []
        self._format(lines, [], 0, len(lines), out)
        res = "\n".join(out)
        return res + "\n" if ends_with_newline and out else res

    # tabwriter.format(): recursive over columns
    def _format(self, lines, widths, line0, line1, out):
        column = len(widths)
        this = line0
        while this < line1:
            line = lines[this]
            if column >= len(line) - 1:
                this += 1
                continue
            self._write_lines(lines, widths, line0, this, out)
            line0 = this
            width = self.minwidth
            while this < line1:
                line = lines[this]
                if column >= len(line) - 1:
                    break
                w = len(line[column]) + self.padding
                if w > width:
                    width = w
                this += 1
            self._format(lines, widths + [width], line0, this, out)
            line0 = this
        self._write_lines(lines, widths, line0, line1, out)

    def _write_lines(self, lines, widths, line0, line1, out):
        for i in range(line0, line1):
            buf = []
            for j, c in enumerate(lines[i]):
                buf.append(c)
                if j < len(widths):
                    buf.append(self.padchar * max(widths[j] - len(c), 0))
            out.append("".join(buf))
