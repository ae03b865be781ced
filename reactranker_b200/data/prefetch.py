"""One-batch look-ahead for training loops that read a result back every step.

The reference's loop (train/train_listwise.py:183-188) plans a batch, featurises it and launches the step strictly in
sequence.  That is free while nothing waits for the GPU: the kernel launches are asynchronous, so the host runs ahead and
prepares batch i+1 while step i executes.  A loop that reads the loss of every step (logging, NaN checks, a benchmark's
device->host read) loses that overlap -- unless batch i+1 is prepared BETWEEN enqueuing step i and reading its result:

    feed = Lookahead(plan, featurise)          # featurise: plan item -> whatever the step needs (graphs already on the device)
    while feed.current is not None:
        loss = step(feed.current)              # enqueue forward / backward / optimiser: returns immediately
        feed.advance()                         # host work + H2D of the NEXT batch, overlapping the GPU
        value = float(loss)                    # only now wait

A worker thread was tried first and was slower: it holds the GIL in 5 ms slices exactly while the launching thread needs it.
"""
from __future__ import annotations

from typing import Callable, Iterable


class Lookahead:
    def __init__(self, batches: Iterable, prepare: Callable):
        self._it = iter(batches)
        self._prepare = prepare
        self.current = None
        self.advance()

    def advance(self) -> None:
        """Prepare the next batch (``current`` becomes None when the plan is exhausted)."""
        try:
            self.current = self._prepare(next(self._it))
        except StopIteration:
            self.current = None
