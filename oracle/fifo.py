"""ORACLE (test infrastructure): the FIFO diagonal-queue index schedule as pure integer arithmetic.

Restates the controller loop of cogvideo_fifo_mp_v2 (longvgen/fifo_sampling/cogvideo_sampling_mp_fifo.py:175-188,
223-259, 322-327, 340-358) and the queue priming of the base stage (pipeline_cogvideox_mp_fifo.py:1190-1194) without any
tensors.  Bit-exact class: tests compare tokensgen_b200.fifo's schedule against this and against tests/golden/fifo_trace_*.json
(traced from the reference controller itself with a recording stub in place of the worker).
"""
from __future__ import annotations

from dataclasses import dataclass, asdict
from typing import List

import numpy as np


@dataclass
class Window:
    iteration: int
    rank: int
    start: int       # first queue slot fed to the DiT (after adaptive-padding clamp)
    mid: int         # first slot written back when the window is not the clamped one
    end: int         # start + nf
    real_end: int    # unclamped start + nf
    write_lo: int    # slots [write_lo, write_hi) of the queue receive this window's outputs
    write_hi: int


def window_schedule(num_frames: int, num_inference_steps: int = 52, nf: int = 13, num_partitions: int = 4,
                    adaptive_padding: bool = True) -> List[List[Window]]:
    l_nf, r_nf = nf - nf // 2, nf // 2
    num_rank = 2 * num_partitions
    queue_start = num_inference_steps - l_nf if adaptive_padding else 0
    out = []
    for it in range(num_frames + num_inference_steps - nf):
        wins = []
        for rank in range(num_rank):
            start = nf * (rank // 2) + r_nf * (rank % 2)
            nxt = nf * ((rank + 1) // 2) + r_nf * ((rank + 1) % 2)
            if nxt <= queue_start:
                continue
            mid = start + (l_nf if rank % 2 == 1 else r_nf)
            real_end = start + nf
            if start < queue_start:
                start = queue_start
            end = start + nf
            if start > queue_start:
                lo, hi = mid, end
            else:
                lo, hi = max(r_nf, start), real_end
            wins.append(Window(it, rank, start, mid, end, real_end, lo, hi))
        out.append(wins)
        queue_start = max(0, queue_start - 1)
    return out


def fifo_timestep_tables(timesteps: np.ndarray, nf: int = 13):
    """:182-185, already flipped the way the controller indexes them (`.flip(0)[start:end]`)."""
    r_nf = nf // 2
    t = np.concatenate([timesteps, np.full(r_nf, timesteps[-1])])
    prev = np.concatenate([timesteps[1:], np.full(r_nf + 1, -1)])
    nxt = np.concatenate([np.full(1, -1), timesteps[:-1], np.full(r_nf, timesteps[-2])])
    return t[::-1].copy(), prev[::-1].copy(), nxt[::-1].copy()


def schedule_as_json(num_frames: int, **kw):
    return [[asdict(w) for w in wins] for wins in window_schedule(num_frames, **kw)]
