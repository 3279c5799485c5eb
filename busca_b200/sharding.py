"""Sequence sharding across the GPUs of one box.

The reference runs one tracker per sequence, sequentially (adapters/ByteTrack/tools/track.py drives one
``BYTETracker`` + one ``BUSCA`` per video, byte_tracker.py:217-221); trackers of different sequences share nothing but
read-only weights.  The hot path therefore shards BY SEQUENCE: one process per GPU, each owning the patch banks of its
sequences, and **no collective on the hot path**.  ``torch.distributed`` (NCCL on the GPU box, gloo in the CPU tests) is
used only after the sequences are done, to gather the per-rank result tables (frame, track id, box, probability - the
rows the MOT writer of mot_evaluator.py:30-40 prints) and the timing maxima.

Nothing here touches CUDA: the functions take the process group as given.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

RESULT_COLS = 8     # seq, frame, track id, x, y, w, h, probability  (float64 rows)


def partition_sequences(n_frames: Sequence[int], world: int, policy: str = "longest_first") -> List[List[int]]:
    """Assign sequence indices to ``world`` ranks.

    ``round_robin``  : sequence i -> rank i % world (static; what the weak-scaling bench uses: equal sequences).
    ``longest_first``: LPT greedy on the frame counts (ties -> lower rank, lower sequence index first), so that the
                       slowest rank - the one the whole job waits for - is as short as a static assignment gets.
    Every sequence appears exactly once; ranks may be empty when there are fewer sequences than ranks."""
    if world < 1:
        raise ValueError("world must be >= 1")
    n = len(n_frames)
    out: List[List[int]] = [[] for _ in range(world)]
    if policy == "round_robin":
        for i in range(n):
            out[i % world].append(i)
        return out
    if policy != "longest_first":
        raise ValueError(f"unknown policy {policy!r}")
    load = [0] * world
    for i in sorted(range(n), key=lambda j: (-int(n_frames[j]), j)):
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += int(n_frames[i])
    for r in out:
        r.sort()
    return out


def makespan(n_frames: Sequence[int], parts: List[List[int]]) -> int:
    return max((sum(int(n_frames[i]) for i in p) for p in parts), default=0)


def pack_results(rows: Sequence[Tuple[int, int, int, float, float, float, float, float]]) -> np.ndarray:
    """Rows (seq, frame, track id, x, y, w, h, prob) -> float64 [n, 8] (ids below 2**53 are exact in float64)."""
    a = np.asarray(list(rows), dtype=np.float64).reshape(-1, RESULT_COLS)
    return a


def gather_results(local: np.ndarray, dist=None, group=None, device: Optional[str] = None) -> Optional[np.ndarray]:
    """Gather ragged per-rank result tables on rank 0, ordered by (sequence, frame, track id).

    Two collectives, both off the hot path: an all-gather of the row counts and an all-gather of the tables padded to
    the maximum count (NCCL has no ragged gather; a few MB at MOT20 scale).  Returns the merged table on rank 0 and
    ``None`` elsewhere; without a process group it just sorts the local table."""
    local = np.ascontiguousarray(local, dtype=np.float64).reshape(-1, RESULT_COLS)
    if dist is None or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return sort_results(local)
    import torch
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = device or "cpu"
    cnt = torch.tensor([local.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(counts, cnt, group=group)
    counts = [int(c.item()) for c in counts]
    mx = max(max(counts), 1)
    buf = torch.zeros((mx, RESULT_COLS), dtype=torch.float64, device=dev)
    if local.shape[0]:
        buf[: local.shape[0]] = torch.from_numpy(local).to(dev)
    parts = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    if rank != 0:
        return None
    merged = np.concatenate([p[:n].cpu().numpy() for p, n in zip(parts, counts)], axis=0) if sum(counts) else np.zeros((0, RESULT_COLS))
    return sort_results(merged)


def sort_results(a: np.ndarray) -> np.ndarray:
    if a.shape[0] == 0:
        return a.reshape(0, RESULT_COLS)
    order = np.lexsort((a[:, 2], a[:, 1], a[:, 0]))
    return a[order]


def reduce_max(values: Sequence[float], dist=None, group=None, device: Optional[str] = None) -> List[float]:
    """Element-wise MAX over ranks of a few scalars (device-timed milliseconds): the job is as slow as its slowest rank."""
    if dist is None or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return [float(v) for v in values]
    import torch
    t = torch.tensor(list(values), dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return [float(v) for v in t.cpu()]


def sequence_seeds(world: int, rank: int, per_rank: int, base: int = 100) -> List[int]:
    """Seeds of the synthetic sequences rank ``rank`` owns in the weak-scaling bench (fixed work per GPU): disjoint
    across ranks and independent of ``world``, so a sequence is the same whether 1 or 8 GPUs run the job."""
    return [base + rank * per_rank + i for i in range(per_rank)]


def write_mot_txt(rows: np.ndarray, seq: int) -> str:
    """MOTChallenge text lines of one sequence, ``frame,id,x1,y1,w,h,s,-1,-1,-1`` with the reference's rounding
    (box to 1 decimal, score to 2, Python float repr; negative ids skipped - mot_evaluator.py:30-40)."""
    out = []
    for r in rows[rows[:, 0] == seq]:
        if r[2] < 0:
            continue
        x, y, w, h = (round(float(v), 1) for v in r[3:7])
        out.append(f"{int(r[1])},{int(r[2])},{x},{y},{w},{h},{round(float(r[7]), 2)},-1,-1,-1\n")
    return "".join(out)
