"""N > 1 path on CPU: two gloo ranks shard one batch contiguously, each checks its slice (the oracle stands in for
the device here), and both end up with the full, correct status array."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from jubjub_schnorr_b200.sharding import shard_range


def test_shard_ranges_cover_exactly():
    for n in (0, 1, 7, 8, 1000, 1 << 20):
        for world in (1, 2, 3, 8):
            cuts = [shard_range(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n, out_dir):
    import torch.distributed as dist
    from tests.sharded_verify import verify_sharded
    from oracle import c_oracle as co
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pk, sig, msg = co.gen_single(77, n, threads=1)
    sig[::5, 0] ^= 1
    calls = []

    def verify(p, s, m):
        calls.append(m.shape[0])
        return co.verify_single(p, s, m, threads=1)[0]

    st = verify_sharded(verify, pk, sig, msg)
    np.save(os.path.join(out_dir, f"st{rank}.npy"), st)
    np.save(os.path.join(out_dir, f"calls{rank}.npy"), np.array(calls))
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_verify(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    n = 101
    mp.spawn(_worker, args=(2, port, n, str(tmp_path)), nprocs=2, join=True)
    from oracle import c_oracle as co
    pk, sig, msg = co.gen_single(77, n, threads=1)
    sig[::5, 0] ^= 1
    expect = co.verify_single(pk, sig, msg)[0]
    for r in range(2):
        assert np.array_equal(np.load(tmp_path / f"st{r}.npy"), expect)
    assert int(np.load(tmp_path / "calls0.npy")[0]) + int(np.load(tmp_path / "calls1.npy")[0]) == n
    assert expect[::5].tolist() == [1] * len(expect[::5])
