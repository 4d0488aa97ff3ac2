"""Ensembles of independent trajectories (SURVEY.md 8d cfg 2, 8e).

Two ways to step a batch ``u0`` of shape ``(B, n_c)``:

* **shared dt** (cfg 2a, the fast path): pass the whole batch to one solver.  The reference's own
  broadcast semantics apply -- one dt for everybody, global max / global 2-norms over the batch
  (solveras.py:451-454) -- and the batch may be sharded over GPUs (``group=``).
* **independent dt** (cfg 2b): every trajectory is its own adaptive problem with its own controller.
  ``evolve_independent`` below runs one plan per trajectory, round-robin on a few CUDA streams so
  that the small kernels of different trajectories overlap.  Results are identical to B separate
  reference runs.  This is the functional form; per-trajectory control blocks inside one set of
  batched kernels (per-row h, roles and coefficient arrays) are not built yet.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch


def evolve_independent(solver_factory: Callable[[], "object"], u0: torch.Tensor, t0: float, tf: float,
                       h_init: Optional[float] = None, streams: int = 8) -> Tuple[torch.Tensor, List[list]]:
    """Evolve every row of ``u0`` with its own dt sequence.

    ``solver_factory()`` must return a fresh adaptive solver (e.g. ``lambda: ETD35(lin_op, nl, config)``).
    Returns the stacked final states and, per trajectory, its trial log ``[(h, s, accepted, t_after), ...]``.
    Trajectories can also be sharded over ranks first (``dist.shard_batch``): there is no collective.
    """
    if u0.dim() < 2:
        raise ValueError("u0 must have a leading batch dimension")
    pool = [torch.cuda.Stream(device=u0.device) for _ in range(max(1, min(streams, u0.shape[0])))]
    main = torch.cuda.current_stream(u0.device)
    out = torch.empty_like(u0, dtype=torch.complex128)
    logs: List[list] = []
    for b in range(u0.shape[0]):
        st = pool[b % len(pool)]
        st.wait_stream(main)
        with torch.cuda.stream(st):
            sol = solver_factory()
            out[b] = sol.evolve(u0[b], t0, tf, h_init=h_init, store_data=False)
            logs.append(list(sol.trial_log))
    for st in pool:
        main.wait_stream(st)
    return out, logs
