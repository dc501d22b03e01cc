"""Multi-GPU host logic on CPU: row-group partitioning and the shard scatter, under gloo with world_size 2 and 3.

The data path itself has no collective (row-groups are independent); what is tested here is that every rank ends up
with a self-contained, correctly rebased shard of the column rank 0 holds — decodable on its own — and that the
shards tile the column exactly.  The CPU checker does the encoding/decoding so that no GPU is needed.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_plan_shards_tiles_whole_rowgroups():
    from alp_b200.shard import plan_shards

    for n_vec in (1, 99, 100, 101, 1234, 1048576, 10486 * 100 - 24):
        for world in (1, 2, 3, 4, 8):
            plan = plan_shards(n_vec, world)
            assert len(plan) == world
            pos = 0
            for first, count in plan:
                assert first == pos or count == 0
                assert first % 100 == 0 or count == 0
                pos = first + count if count else pos
            assert pos == n_vec
            counts = [-(-c // 100) for _, c in plan]
            assert max(counts) - min(counts) <= 1  # balanced to within one row-group


def _worker(rank, world, port, kind, n_values, out_dir):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from alp_b200 import shard
    from oracle import pyoracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    checker = pyoracle.port()
    x = pyoracle.generate(n_values, kind)  # every rank can regenerate the column: stateless generator
    vb = x.dtype.itemsize
    col = shard.host_column_tensors(checker.encode_column(x)) if rank == 0 else None
    mine, (first, count) = shard.scatter_column(col, src=0, value_bytes=vb)
    assert mine["meta"].shape[0] == count
    h = shard.tensors_to_host_column(mine, vb)
    ok = True
    if count:
        dec = checker.decode_column(h)
        ok = dec.tobytes() == x[first * 1024 : (first + count) * 1024].tobytes()
        # the shard is self-contained: offsets start at zero and are dense
        ok = ok and int(h.meta["packed_off"][0]) == 0 and int(h.meta["exc_off"][0]) == 0
    flags = torch.tensor([1 if ok else 0, count], dtype=torch.int64)
    gathered = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, flags)
    if rank == 0:
        total = sum(int(g[1]) for g in gathered)
        good = all(int(g[0]) == 1 for g in gathered) and total == n_values // 1024
        with open(os.path.join(out_dir, "result_%d_%d" % (kind, world)), "w") as fh:
            fh.write("ok" if good else "bad %s" % [g.tolist() for g in gathered])
    dist.destroy_process_group()


@pytest.mark.parametrize("world,kind", [(2, 2), (2, 3), (3, 4)])
def test_scatter_shards_decode_independently(world, kind, tmp_path):
    n_values = 1024 * 437  # 4 full row-groups + a short one
    mp.spawn(_worker, args=(world, _free_port(), kind, n_values, str(tmp_path)), nprocs=world, join=True)
    assert open(os.path.join(str(tmp_path), "result_%d_%d" % (kind, world))).read() == "ok"


def test_slice_column_rebases_offsets(port):
    from alp_b200 import shard
    from oracle import pyoracle

    x = np.concatenate([pyoracle.generate(102400, 2), pyoracle.generate(102400, 3)])
    t = shard.host_column_tensors(port.encode_column(x))
    for first, count in ((0, 100), (100, 100), (37, 120), (199, 1), (0, 0)):
        s = shard.slice_column(t, first, count)
        h = shard.tensors_to_host_column(s, 8)
        assert h.n_vectors == count
        if count:
            assert port.decode_column(h).tobytes() == x[first * 1024 : (first + count) * 1024].tobytes()


def test_slice_column_of_a_completion_order_layout(port):
    """Shards of a column whose blocks are NOT in vector order (alpb200_encode_unordered_*; here: a random permutation of
    an ordered column's blocks) still decode on their own: slice_column takes the min/max byte range of the records."""
    from conftest import shuffle_layout

    from alp_b200 import shard
    from oracle import pyoracle

    x = np.concatenate([pyoracle.generate(102400, 2), pyoracle.generate(102400, 3), pyoracle.generate(30 * 1024, 2, first_index=5)])
    shuffled = shuffle_layout(port.encode_column(x), np.random.default_rng(9))
    assert port.decode_column(shuffled).tobytes() == x.tobytes()
    t = shard.host_column_tensors(shuffled)
    for first, count in ((0, 100), (100, 100), (37, 120), (229, 1), (0, 230)):
        s = shard.slice_column(t, first, count)
        h = shard.tensors_to_host_column(s, 8)
        assert port.decode_column(h).tobytes() == x[first * 1024 : (first + count) * 1024].tobytes()


def _nccl_worker(rank, world, port, kind, n_values, out_dir):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import alp_b200
    from alp_b200 import shard

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    vb = 4 if kind == 4 else 8
    tensors = None
    if rank == 0:  # the column is encoded on GPU 0 and handed out from there
        col = alp_b200.encode(alp_b200.generate(n_values, kind, dev))
        tensors = shard.column_tensors(col)
    mine, (first, count) = shard.scatter_column(tensors, src=0, value_bytes=vb, device=dev)
    ok = mine["meta"].shape[0] == count
    if count:
        got = alp_b200.decode(shard.tensors_to_device_column(mine, vb, dev))
        want = alp_b200.generate(count * 1024, kind, dev, first_index=first * 1024)  # stateless generator: any rank can re-create its slice
        ibits = torch.int64 if vb == 8 else torch.int32
        ok = ok and bool(torch.equal(got.view(ibits), want.view(ibits)))
    flags = torch.tensor([1 if ok else 0, count], dtype=torch.int64, device=dev)
    gathered = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(gathered, flags)
    if rank == 0:
        good = all(int(g[0]) == 1 for g in gathered) and sum(int(g[1]) for g in gathered) == n_values // 1024
        with open(os.path.join(out_dir, "nccl_%d_%d" % (kind, world)), "w") as fh:
            fh.write("ok" if good else "bad %s" % [g.tolist() for g in gathered])
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("kind", [2, 3, 4])
def test_nccl_scatter_shards_decode_on_their_gpus(kind, tmp_path):
    """BASELINE config 5 ("scattered across the GPUs"): a column encoded on GPU 0 goes out as whole-row-group shards over NCCL
    point-to-point sends and every GPU decodes its shard bit-exactly.  Needs two devices (skipped otherwise)."""
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least two CUDA devices")
    n_values = 1024 * 937  # 9 full row-groups + a short one
    mp.spawn(_nccl_worker, args=(world, _free_port(), kind, n_values, str(tmp_path)), nprocs=world, join=True)
    assert open(os.path.join(str(tmp_path), "nccl_%d_%d" % (kind, world))).read() == "ok"
