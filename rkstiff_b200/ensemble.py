"""Ensembles of independent trajectories (SURVEY.md 8d cfg 2, 8e).

Two ways to step a batch ``u0`` of shape ``(B, n_c)``:

* **shared dt** (cfg 2a, the fast path): pass the whole batch to one solver.  The reference's own
  broadcast semantics apply -- one dt for everybody, global max / global 2-norms over the batch
  (solveras.py:451-454) -- and the batch may be sharded over GPUs (``group=``).
* **independent dt** (cfg 2b): every trajectory is its own adaptive problem with its own controller.
  With a fused nonlinearity use ``solver.evolve_independent(u0, t0, tf)`` (solveras.py): one plan
  holds a control block, coefficient arrays and buffer roles per trajectory, and one set of launches
  (``gridDim.z`` = trajectory) steps all of them.  ``evolve_independent`` below is the general form
  for arbitrary torch ``nl_func`` callables: one plan per trajectory, round-robin on a few CUDA
  streams so that the small kernels of different trajectories overlap.  Both give results identical
  to B separate reference runs.
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
