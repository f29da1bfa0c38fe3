"""SIGQUIT stack dump — mirror of pkg/gpu/nvidia/coredump.go (goroutine dump -> thread dump)."""
import sys
import threading
import traceback


def StackTrace(all_: bool = True) -> str:  # coredump.go:8-25
    frames = sys._current_frames()
    out = []
    for t in threading.enumerate():
        f = frames.get(t.ident)
        if f is None:
            continue
        out.append(f"thread {t.name} [{'daemon' if t.daemon else 'main'}]:\n" + "".join(traceback.format_stack(f)))
        if not all_:
            break
    return "\n".join(out)


def coredump(fileName: str) -> None:  # coredump.go:27-30
    with open(fileName, "w") as f:
        f.write(StackTrace(True))
