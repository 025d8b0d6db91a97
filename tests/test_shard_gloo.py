"""The N > 1 host logic on CPU: two gloo ranks shard a channel set, each decodes its own block
(with the oracle standing in for the GPU, this is a plumbing test), and the summed statistics and
the concatenated decisions equal the single-rank run.  No collective touches the data path."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp


def test_partition_covers_everything_once():
    from qpsk_b200.shard import partition
    for n in (1, 7, 64, 65536, 1_000_003):
        for world in (1, 2, 3, 4, 8):
            spans = [partition(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (s0, c0), (s1, _) in zip(spans[:-1], spans[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        partition(10, 2, 2)


def _worker(rank, world, port, pcm_path, out_dir):
    import torch.distributed as dist
    import oracle
    from qpsk_b200.shard import max_over_ranks, partition, reduce_stats
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pcm = np.load(pcm_path)
    start, count = partition(pcm.shape[0], world, rank)
    o = oracle.Oracle()
    out = o.rx_run(pcm[start:start + count], want=("dibit", "freq"))
    stats = reduce_stats([out["dibit"].size, float(np.abs(out["freq"][:, -1].astype(np.float64)).sum()), count])
    worst = max_over_ranks(10.0 + rank)
    np.save(os.path.join(out_dir, "dibit_%d.npy" % rank), out["dibit"])
    if rank == 0:
        np.save(os.path.join(out_dir, "stats.npy"), np.array(stats + [worst]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_rank(oracle_lib, tmp_path):
    from synth import make_pcm
    o = oracle_lib.Oracle()
    pcm, _ = make_pcm(5, 6, seed=8, esn0_db=20.0, oracle=o)
    single = o.rx_run(pcm, want=("dibit", "freq"))
    np.save(tmp_path / "pcm.npy", pcm)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path / "pcm.npy"), str(tmp_path)), nprocs=2, join=True)
    both = np.concatenate([np.load(tmp_path / "dibit_0.npy"), np.load(tmp_path / "dibit_1.npy")])
    assert np.array_equal(both, single["dibit"])
    stats = np.load(tmp_path / "stats.npy")
    assert stats[0] == single["dibit"].size and stats[2] == 5 and stats[3] == 11.0
    assert stats[1] == pytest.approx(float(np.abs(single["freq"][:, -1].astype(np.float64)).sum()), rel=1e-12)
