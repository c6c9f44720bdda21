"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: contiguous sharding, unequal-shard gather, and the
invariance contract of `sample_sharded` (global cell offsets) using a stand-in LatentDiffusion whose rows are a
deterministic function of the global cell index (the CUDA path itself is covered by the -m gpu tests)."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from scldm_b200.dist import gather_rows, sample_sharded, shard_range


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 64, 1001):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


class FakeLDM:
    """rows depend only on (global cell index, label) -- what the Philox keying guarantees on the GPU."""

    def __init__(self):
        self.cells_generated = 0

    def sample(self, condition, guidance_weight, batch_size, genes, cell_offset=None, **kw):
        off = self.cells_generated if cell_offset is None else cell_offset
        self.cells_generated = off + batch_size
        idx = torch.arange(off, off + batch_size, dtype=torch.float32)
        lab = condition["c"].float()
        G = genes.shape[1]
        uncond = idx[:, None] * 10 + torch.arange(G)[None, :]
        guided = uncond + 1000 * (1 + lab[:, None])
        counts = torch.cat([uncond, guided])
        z = torch.cat([idx, idx + 0.5])[:, None, None].expand(-1, 2, 2).contiguous()
        return counts, z


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        a, b = shard_range(n, rank, world)
        local = torch.arange(a, b, dtype=torch.float32)[:, None].repeat(1, 3)
        full = gather_rows(local, n)
        assert torch.equal(full, torch.arange(n, dtype=torch.float32)[:, None].repeat(1, 3))
        cond = {"c": torch.arange(n) % 4}
        genes = torch.arange(1, 6).unsqueeze(0).repeat(n, 1)
        counts, z = sample_sharded(FakeLDM(), cond, {"c": 1.0}, n, genes)
        ref_counts, ref_z = FakeLDM().sample(cond, {"c": 1.0}, n, genes)
        assert torch.equal(counts, ref_counts) and torch.equal(z, ref_z)
        local_counts, _ = sample_sharded(FakeLDM(), cond, {"c": 1.0}, n, genes, gather=False)
        assert local_counts.shape[0] == 2 * (b - a)
        ret[rank] = True
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [7, 64])
def test_two_rank_gather_and_invariance(n):
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, n, ret), nprocs=world, join=True)
        assert all(ret.get(r) for r in range(world))
