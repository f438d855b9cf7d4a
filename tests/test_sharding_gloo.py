"""world_size-2 gloo test of the multi-GPU host logic (no GPU): shard partition, weight broadcast,
root-statistics gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from stochastic_muzero_b200.sharding import broadcast_weights, gather_roots, shard_range
    from stochastic_muzero_b200.weights import ModelShape, random_blob
    shape = ModelShape(4, 2, 2, 61, 126, 4)
    blob = torch.from_numpy(random_blob(shape, seed=5)) if rank == 0 else None
    got = broadcast_weights(blob, shape, src=0)
    lo, hi = shard_range(total, rank, world)
    local = {"visits": torch.arange(lo, hi, dtype=torch.int32)[:, None].repeat(1, 2),
             "root_values": torch.arange(lo, hi, dtype=torch.float32)}
    full = gather_roots(local, total)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), blob=got.numpy(), visits=full["visits"].numpy(),
             values=full["root_values"].numpy(), lo=lo, hi=hi)
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7])
def test_two_rank_sharding_broadcast_and_gather(tmp_path, total):
    from stochastic_muzero_b200.weights import ModelShape, random_blob
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, total, str(tmp_path)), nprocs=world, join=True)
    ref = random_blob(ModelShape(4, 2, 2, 61, 126, 4), seed=5)
    covered = []
    for r in range(world):
        z = np.load(tmp_path / f"r{r}.npz")
        assert np.array_equal(z["blob"], ref), "weight broadcast changed the blob"
        assert np.array_equal(z["visits"][:, 0], np.arange(total)) and np.array_equal(z["values"], np.arange(total))
        covered += list(range(int(z["lo"]), int(z["hi"])))
    assert covered == list(range(total)), "shards must partition the trees exactly"


def test_shard_range_properties():
    from stochastic_muzero_b200.sharding import shard_range
    for total in (1, 5, 4096, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
