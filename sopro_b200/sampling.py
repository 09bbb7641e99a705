"""Host-side sampling helpers of the product path.

The reference draws its randomness inside ``torch.multinomial`` from the global CPU
generator, one call per frame (reference sampling.py:83,93).  ATen implements a
single-sample multinomial as ``argmax(p / q)`` with ``q = empty_like(p).exponential_(1)``,
so the whole random stream of an utterance is a [steps, V] tensor of Exp(1) draws that can be
produced up front and handed to the device sampler: ``noise_tape``.  Because top-k zeroes all
but the ``top_k`` best-ranked probabilities and the draw is indexed by sorted rank
(sampling.py:83-84), only the first ``top_k`` columns are ever needed on the device."""
from __future__ import annotations

from typing import Optional

import torch


def noise_tape(steps: int, vocab: int, *, seed: Optional[int] = None, generator: Optional[torch.Generator] = None,
               keep: Optional[int] = None, pin: bool = False) -> torch.Tensor:
    """[steps, keep or vocab] Exp(1) draws, consumed exactly as `steps` successive
    ``torch.multinomial(p[1, vocab], 1)`` calls would.

    seed=None and generator=None -> the global CPU generator is consumed, which is what the
    reference API does (it has no seed kwarg; the CLI seeds globally, cli.py:72-75)."""
    if seed is not None:
        generator = torch.Generator().manual_seed(int(seed))
    full = torch.empty(int(steps), int(vocab))
    full.exponential_(1.0, generator=generator)
    out = full if keep is None else full[:, : int(keep)].contiguous()
    return out.pin_memory() if pin else out
