"""Index batching with the reference's RNG stream.

The reference iterates ``DataLoader(range(n), batch_size, shuffle=True)`` (train_teacher_gnn.py:35,
main.py:72-73,167-168).  Collating 65,536 Python ints per batch costs more host time than the whole GPU step,
so ``shuffled_batches`` reproduces the exact index stream without the Python lists: DataLoader first draws its
base seed from the global torch generator, then ``RandomSampler`` draws the permutation seed and calls
``torch.randperm(n, generator=g)``.  ``tests/test_host_logic.py`` pins the equality against the real DataLoader.
"""
from __future__ import annotations

from typing import Iterator

import torch


def shuffled_batches(n: int, batch_size: int) -> Iterator[torch.Tensor]:
    """Same index tensors, in the same order and consuming the same global-RNG draws, as
    ``DataLoader(range(n), batch_size, shuffle=True)`` (CPU int64)."""
    torch.empty((), dtype=torch.int64).random_()  # _BaseDataLoaderIter._base_seed
    seed = int(torch.empty((), dtype=torch.int64).random_().item())  # RandomSampler.__iter__
    g = torch.Generator()
    g.manual_seed(seed)
    perm = torch.randperm(n, generator=g)
    for s in range(0, n, batch_size):
        yield perm[s:s + batch_size]


def sequential_batches(n: int, batch_size: int) -> Iterator[torch.Tensor]:
    """``DataLoader(range(n), batch_size)`` without shuffling (the scoring loops, train_teacher_gnn.py:95)."""
    idx = torch.arange(n)
    for s in range(0, n, batch_size):
        yield idx[s:s + batch_size]
