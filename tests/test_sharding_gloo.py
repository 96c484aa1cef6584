"""N>1 host logic on CPU: two `gloo` ranks each own a contiguous shard (seeds and y0 from the GLOBAL index), integrate
it (here with the oracle standing in for the device — the product's kernel is covered by the -m gpu tests), and
all-reduce the moment accumulators; the result must equal the single-process moments and must not depend on the
number of ranks."""

import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from jitcsde_b200 import models, sharding
from oracle import port

TOTAL, T = 48, 0.5
SPEC = models.ou


def _free_port():
	with socket.socket() as s:
		s.bind(("127.0.0.1", 0))
		return s.getsockname()[1]


def _worker(rank, world, port_no, out):
	os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
	dist.init_process_group("gloo", rank=rank, world_size=world)
	lo, hi = sharding.shard_range(TOTAL, rank, world)
	y0 = sharding.global_initial_states(lo, hi, SPEC.n)
	res = port.run_ensemble(SPEC, y0, sharding.global_seeds(lo, hi), 0.0, T)
	shift = np.array([0.1, 0.9])
	acc = torch.from_numpy(sharding.accumulate_moments(res["y"], shift))
	sharding.all_reduce_sum(acc)
	if rank == 0:
		np.save(out, acc.numpy())
	dist.destroy_process_group()


def test_shard_ranges_partition_the_ensemble():
	for total in (1, 7, 48, 2**20):
		for world in (1, 2, 3, 8):
			r = [sharding.shard_range(total, k, world) for k in range(world)]
			assert r[0][0] == 0 and r[-1][1] == total
			assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))


def test_inputs_depend_on_the_global_index_only():
	full = sharding.global_initial_states(0, 10, 3)
	assert (sharding.global_initial_states(4, 10, 3) == full[4:]).all()
	assert (sharding.global_seeds(2**32 - 2, 2**32 + 2) == np.array([2**32 - 2, 2**32 - 1, 0, 1], dtype=np.uint32)).all()


def test_two_rank_moments_equal_single_process(tmp_path):
	out = str(tmp_path / "acc.npy")
	mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
	acc2 = np.load(out)
	y0 = sharding.global_initial_states(0, TOTAL, SPEC.n)
	res = port.run_ensemble(SPEC, y0, sharding.global_seeds(0, TOTAL), 0.0, T)
	shift = np.array([0.1, 0.9])
	acc1 = sharding.accumulate_moments(res["y"], shift)
	np.testing.assert_allclose(acc2, acc1, rtol=1e-13)
	m = sharding.moments_from_accumulators(acc2, shift)
	np.testing.assert_allclose(m["mean"], res["y"].mean(axis=0), rtol=1e-12)
	np.testing.assert_allclose(m["var"], res["y"].var(axis=0, ddof=1), rtol=1e-10)
