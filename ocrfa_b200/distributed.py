"""Data-parallel driver of the recognition path: one process per GPU, lines sharded contiguously, no collective on
the compute path. The reference's analogue is `torch.nn.DataParallel(model)` inside get_recognizer
(/root/reference/src/ocrfa/utils.py:692), a single-process scatter/gather over the batch dimension.

Every text line is independent (eval-mode BatchNorm, no cross-line state in the model or the decode), so rank r takes
lines [r*ceil(N/R), (r+1)*ceil(N/R)) and the only communication is an all_gather of the fixed-width decoded label
tensors, their lengths and the confidences AFTER the compute (about 0.5 KB per line). Backend: NCCL on GPUs; the
same code runs over gloo on CPU tensors, which is how the sharding/gather logic is tested without GPUs.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_lines: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of rank `rank`: ceil(N/R) lines each, the tail ranks may be short or empty."""
    if n_lines < 0 or world <= 0 or not (0 <= rank < world):
        raise ValueError("bad shard arguments")
    per = (n_lines + world - 1) // world
    lo = min(rank * per, n_lines)
    return lo, min(lo + per, n_lines)


def gather_decoded(labels: torch.Tensor, lengths: torch.Tensor, conf: torch.Tensor, n_lines: int,
                   group: Optional[dist.ProcessGroup] = None):
    """all_gather of per-rank results (labels int32 [n_r, T], lengths int32 [n_r], conf float64 [n_r]) into the
    global line order. Every rank returns the full (labels [N,T], lengths [N], conf [N]) on its own device."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return labels[:n_lines], lengths[:n_lines], conf[:n_lines]
    world = dist.get_world_size(group)
    per = (n_lines + world - 1) // world
    T = labels.shape[1]
    dev = labels.device

    def padded(t, shape, dtype):
        out = torch.zeros(shape, dtype=dtype, device=dev)
        out[: t.shape[0]] = t
        return out

    lab = padded(labels, (per, T), torch.int32)
    ln = padded(lengths, (per,), torch.int32)
    cf = padded(conf, (per,), torch.float64)
    labs = [torch.empty_like(lab) for _ in range(world)]
    lns = [torch.empty_like(ln) for _ in range(world)]
    cfs = [torch.empty_like(cf) for _ in range(world)]
    dist.all_gather(labs, lab, group=group)
    dist.all_gather(lns, ln, group=group)
    dist.all_gather(cfs, cf, group=group)
    return torch.cat(labs)[:n_lines], torch.cat(lns)[:n_lines], torch.cat(cfs)[:n_lines]


def recognize_sharded(n_lines: int, recognize_batch: Callable[[int, int], Tuple[torch.Tensor, torch.Tensor, torch.Tensor]],
                      t_steps: int, batch_size: int = 256, device: Optional[torch.device] = None,
                      group: Optional[dist.ProcessGroup] = None):
    """Runs `recognize_batch(lo, hi)` (-> labels [hi-lo, T] int32, lengths int32, conf float64, on `device`) over this
    rank's shard in inner batches of `batch_size`, then gathers the global result on every rank."""
    rank = dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    lo, hi = shard_range(n_lines, rank, world)
    dev = device or torch.device("cpu")
    labs, lns, cfs = [], [], []
    for b0 in range(lo, hi, batch_size):
        la, ln, cf = recognize_batch(b0, min(b0 + batch_size, hi))
        labs.append(la.to(dev, torch.int32))
        lns.append(ln.to(dev, torch.int32))
        cfs.append(cf.to(dev, torch.float64))
    if labs:
        la, ln, cf = torch.cat(labs), torch.cat(lns), torch.cat(cfs)
    else:
        la = torch.zeros((0, t_steps), dtype=torch.int32, device=dev)
        ln = torch.zeros((0,), dtype=torch.int32, device=dev)
        cf = torch.zeros((0,), dtype=torch.float64, device=dev)
    return gather_decoded(la, ln, cf, n_lines, group)


def labels_to_strings(labels: torch.Tensor, lengths: torch.Tensor, character: Sequence[str]) -> List[str]:
    lab = labels.cpu().numpy()
    ln = lengths.cpu().numpy()
    return ["".join(character[k] for k in row[:n]) for row, n in zip(lab.tolist(), ln.tolist())]


def max_over_ranks(value: float, device: torch.device, group: Optional[dist.ProcessGroup] = None) -> float:
    """Device time of the slowest rank (the multi-GPU metric is N_lines / max_r time_r)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
