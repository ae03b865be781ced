"""Host-side batch prefetch: plan + featurise batch i+1 on a worker thread while the GPU runs step i.

The reference's training loop (train/train_listwise.py:183-188) plans a batch, featurises it and launches the step strictly in
sequence.  That is free when nothing waits for the GPU; as soon as the caller reads a step's result (a loss for logging, NaN
checks, a benchmark's device->host read) the host work of the next step is exposed.  ``prefetch_batches`` keeps ``depth``
prepared batches ahead.  pandas / numpy release the GIL in their hot loops and ctypes releases it for the kernel launches, so
the worker overlaps with the launching thread."""
from __future__ import annotations

import queue
import threading
from typing import Callable, Iterable, Iterator


class _Failure:
    def __init__(self, exc: BaseException):
        self.exc = exc


_DONE = object()


def prefetch_batches(batches: Iterable, prepare: Callable, depth: int = 2) -> Iterator:
    """Yield ``prepare(batch)`` for every ``batch`` of ``batches``, computed up to ``depth`` items ahead on a worker thread.
    Exceptions of the worker are re-raised at the consumer; abandoning the generator stops the worker."""
    q: "queue.Queue" = queue.Queue(maxsize=max(1, depth))
    stop = threading.Event()

    def work():
        try:
            for b in batches:
                item = prepare(b)
                while not stop.is_set():
                    try:
                        q.put(item, timeout=0.1)
                        break
                    except queue.Full:
                        continue
                if stop.is_set():
                    return
            q.put(_DONE)
        except BaseException as e:  # noqa: BLE001 - handed to the consumer
            q.put(_Failure(e))

    t = threading.Thread(target=work, name="rr-batch-prefetch", daemon=True)
    t.start()
    try:
        while True:
            item = q.get()
            if item is _DONE:
                return
            if isinstance(item, _Failure):
                raise item.exc
            yield item
    finally:
        stop.set()
