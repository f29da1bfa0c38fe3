"""FS + signal watchers — mirror of pkg/gpu/nvidia/watchers.go (fsnotify -> inotify(7) via libc)."""
from __future__ import annotations

import ctypes
import ctypes.util
import os
import queue
import select
import signal
import struct
import threading
from typing import List, Tuple

IN_CREATE = 0x00000100
IN_DELETE = 0x00000200
_EVENT = struct.Struct("iIII")


class FSWatcher:
    """fsnotify.Watcher on directories: `Events` yields (path, mask) like fsnotify's Event{Name, Op}."""

    def __init__(self, *files: str):
        self._libc = ctypes.CDLL(ctypes.util.find_library("c") or "libc.so.6", use_errno=True)
        self._fd = self._libc.inotify_init1(os.O_NONBLOCK | os.O_CLOEXEC)
        if self._fd < 0:
            raise OSError(ctypes.get_errno(), "inotify_init1")
        self._wd = {}
        self.Events: "queue.Queue[Tuple[str, int]]" = queue.Queue()
        self.Errors: "queue.Queue[Exception]" = queue.Queue()
        for f in files:
            wd = self._libc.inotify_add_watch(self._fd, f.encode(), IN_CREATE | IN_DELETE)
            if wd < 0:
                err = ctypes.get_errno()
                self.Close()
                raise OSError(err, f"inotify_add_watch({f})")
            self._wd[wd] = f
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, name="fswatcher", daemon=True)
        self._t.start()

    def _run(self):
        while not self._stop.is_set():
            r, _, _ = select.select([self._fd], [], [], 0.2)
            if not r:
                continue
            try:
                buf = os.read(self._fd, 65536)
            except BlockingIOError:
                continue
            except OSError as e:
                self.Errors.put(e)
                return
            off = 0
            while off + _EVENT.size <= len(buf):
                wd, mask, _cookie, ln = _EVENT.unpack_from(buf, off)
                name = buf[off + _EVENT.size: off + _EVENT.size + ln].split(b"\0", 1)[0].decode()
                off += _EVENT.size + ln
                self.Events.put((os.path.join(self._wd.get(wd, ""), name), mask))

    def Close(self):
        if getattr(self, "_stop", None):
            self._stop.set()
        if self._fd >= 0:
            os.close(self._fd)
            self._fd = -1


def newFSWatcher(*files: str) -> FSWatcher:  # watchers.go:10-25
    return FSWatcher(*files)


def newOSWatcher(*sigs: int) -> "queue.Queue[int]":  # watchers.go:27-32 (main thread only, like signal.Notify)
    q: "queue.Queue[int]" = queue.Queue(maxsize=16)
    for s in sigs:
        signal.signal(s, lambda signum, _frame: q.put_nowait(signum))
    return q
