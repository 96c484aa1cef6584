"""
Multi-GPU layout of an ensemble (SURVEY.md §8 e): trajectories are independent, so rank r of G owns the contiguous
block of global indices [r*B/G, (r+1)*B/G) — its own states, generator streams and noise memories — and nothing is
exchanged during integration.  The only collective is the all-reduce (sum) of the 1+2n moment accumulators
{count, Σ(y−c), Σ(y−c)²} at the end.  Initial states and seeds are functions of the GLOBAL index, so every
trajectory's result is independent of G.
"""

from __future__ import annotations

import numpy as np


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
	"""Contiguous block of global trajectory indices owned by `rank` (first ranks get the remainder)."""
	base, extra = divmod(total, world)
	lo = rank * base + min(rank, extra)
	return lo, lo + base + (1 if rank < extra else 0)


def global_seeds(lo: int, hi: int) -> np.ndarray:
	"""seed_k = k mod 2^32 (SURVEY.md §8 d)"""
	return (np.arange(lo, hi, dtype=np.uint64) & np.uint64(0xFFFFFFFF)).astype(np.uint32)


def global_initial_states(lo: int, hi: int, n: int, seed: int = 42) -> np.ndarray:
	"""rows lo..hi-1 of default_rng(seed).random((hi, n)) — the y0 rule of the benchmark workload"""
	return np.random.default_rng(seed).random((hi, n))[lo:]


def accumulate_moments(y: np.ndarray, shift: np.ndarray) -> np.ndarray:
	"""Host restatement of the device accumulator layout (used by tests): [count, Σ(y−c), Σ(y−c)²]"""
	d = y - shift
	return np.concatenate([[float(len(y))], d.sum(axis=0), (d * d).sum(axis=0)])


def moments_from_accumulators(acc: np.ndarray, shift: np.ndarray) -> dict:
	n = (len(acc) - 1) // 2
	count = acc[0]
	mean_c = acc[1:1 + n] / count
	var = (acc[1 + n:] - count * mean_c ** 2) / (count - 1) if count > 1 else np.full(n, np.nan)
	return {"count": int(count), "mean": mean_c + shift, "var": var}


def all_reduce_sum(acc_tensor, group=None):
	"""Sum the accumulators over all ranks (NCCL on GPUs, gloo in the CPU tests); no-op without a process group."""
	import torch.distributed as dist
	if dist.is_available() and dist.is_initialized():
		dist.all_reduce(acc_tensor, op=dist.ReduceOp.SUM, group=group)
	return acc_tensor
