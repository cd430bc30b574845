"""world_size-2 gloo run of the multi-GPU plumbing (fan-out from the ingest rank, max-over-ranks timing,
result gather) on CPU. The per-channel processing function is injected: here the CPU oracle stands in for
the CUDA chain, which is what the -m gpu tests and bench.py --gpus N exercise on real devices."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gr_air_modes_b200 import shard, synth


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port_no, n, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import cpu_oracle as co
    dev = torch.device("cpu")

    def make(ch):
        sc = synth.make_scene(4e6, n, 12, seed=100 + ch)
        return torch.from_numpy(sc.iq.copy()), [b.frame.hex() for b in sc.bursts]

    iq, sent = shard.fan_out(make, rank, world, dev, 2 * n)
    expect = synth.make_scene(4e6, n, 12, seed=100 + rank)
    assert np.array_equal(iq.numpy(), expect.iq)                  # each rank got ITS channel, bit for bit
    assert sent == [b.frame.hex() for b in expect.bursts]
    msgs = co.Port().run_iq(iq.numpy(), 4e6, 7.0, True, co.MA_SLIDING64).msgs
    counts = shard.gather_counts(len(msgs), world, dev)
    slowest = shard.max_over_ranks(1.0 + rank, world, dev)
    assert slowest == float(world)
    assert counts[rank] == len(msgs) and len(counts) == world
    ret[rank] = counts
    dist.barrier()
    dist.destroy_process_group()


def test_fan_out_and_gather_world2():
    world, n = 2, 120_000
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n, ret), nprocs=world, join=True)
    assert ret[0] == ret[1] and all(c > 0 for c in ret[0])


def test_channel_ownership():
    assert shard.channels_of(1, 4, 8) == [1, 5]
    assert sorted(sum((shard.channels_of(r, 3, 8) for r in range(3)), [])) == list(range(8))
    assert shard.max_over_ranks(3.5, 1, None) == 3.5 and shard.gather_counts(7, 1, None) == [7]
