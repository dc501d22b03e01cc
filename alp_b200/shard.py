"""Multi-GPU: partition a compressed column by whole row-groups and scatter the shards.

The codec has no exchange step — every row-group is encoded and decoded independently (reference
include/alp/encoder.hpp:420-427: all encoder state is per row-group) — so N GPUs simply own disjoint, contiguous
ranges of row-groups and there is NO collective on the data path.  The only communication is the one-off hand-out of
compressed shards from the rank that holds the column, a plain scatter of independent byte ranges
(torch.distributed send/recv: NCCL over NVLink on GPUs, gloo on CPU for the tests).  Keep it outside any timed decode
region: one GPU's egress is ~0.8 TB/s, a decode pass moves 5+ TB/s.

Works on tensors of any device: the CPU tests run it under gloo with world_size 2.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import _abi


def plan_shards(n_vectors, world_size):
    """Vector ranges [first, first+count) per rank: whole row-groups (100 vectors), as even as possible, in order.

    Every rank gets a contiguous range; ranks at the end may get an empty range when there are fewer row-groups than
    ranks.  The last row-group may be short (n_vectors % 100)."""
    n_rg = -(-n_vectors // _abi.ROWGROUP_VECTORS)
    base, extra = divmod(n_rg, world_size)
    out, rg = [], 0
    for r in range(world_size):
        take = base + (1 if r < extra else 0)
        first = rg * _abi.ROWGROUP_VECTORS
        last = min(n_vectors, (rg + take) * _abi.ROWGROUP_VECTORS)
        out.append((first, max(0, last - first)))
        rg += take
    return out


def _units(meta):
    """packed block size of every record in 128-byte units (meta: uint8 [n, 32] tensor on any device)."""
    scheme = meta[:, 26].to(torch.int64)
    bw = meta[:, 27].to(torch.int64)
    e = meta[:, 28].to(torch.int64)
    return torch.where(scheme == _abi.SCHEME_ALP_RD, bw + e, bw)


def _u32(meta, byte):
    return meta[:, byte : byte + 4].contiguous().view(torch.int32).to(torch.int64).reshape(-1) & 0xFFFFFFFF


def slice_column(col, first, count):
    """Cut vectors [first, first+count) out of a column given as a dict of tensors
    {meta uint8 [n,32], packed uint8, exc_val, exc_pos int16} and rebase the offsets in the metadata records.

    The byte range is the min / max over the records of the range: for a vector-order column that is exactly the
    range's own bytes; for a completion-order column (encode(ordered=False)) it also carries the few foreign blocks
    that finished in between — harmless, the shard's records never point at them."""
    meta = col["meta"][first : first + count].clone()
    if count == 0:
        return {"meta": meta, "packed": col["packed"][:0].clone(), "exc_val": col["exc_val"][:0].clone(), "exc_pos": col["exc_pos"][:0].clone()}
    poff, eoff = _u32(meta, 16), _u32(meta, 20)
    cnt = meta[:, 24:26].contiguous().view(torch.int16).to(torch.int64).reshape(-1) & 0xFFFF
    p0, e0 = int(poff.min()), int(eoff.min())
    p1 = int((poff + _units(meta)).max())
    e1 = int((eoff + cnt).max())
    meta[:, 16:20] = (poff - p0).to(torch.int32).view(torch.uint8).reshape(-1, 4)
    meta[:, 20:24] = (eoff - e0).to(torch.int32).view(torch.uint8).reshape(-1, 4)
    return {
        "meta": meta,
        "packed": col["packed"][p0 * 128 : p1 * 128].clone(),
        "exc_val": col["exc_val"][e0:e1].clone(),
        "exc_pos": col["exc_pos"][e0:e1].clone(),
    }


def scatter_column(col, src=0, group=None, value_bytes=8, device=None):
    """Hand every rank its shard of the column held by rank `src`.

    col: on `src` a dict of tensors as in slice_column (other ranks pass None).  Returns (shard dict, (first, count)).
    Sizes travel first (one broadcast), then each shard's four arrays (point-to-point)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    device = device or (col["meta"].device if col is not None else torch.device("cpu"))
    header = torch.zeros(1 + 4 * world, dtype=torch.int64, device=device)
    shards = None
    if rank == src:
        n_vec = col["meta"].shape[0]
        plan = plan_shards(n_vec, world)
        shards = [slice_column(col, f, c) for f, c in plan]
        header[0] = n_vec
        for r, s in enumerate(shards):
            header[1 + 4 * r : 5 + 4 * r] = torch.tensor([s["meta"].shape[0], s["packed"].numel(), s["exc_val"].numel(), s["exc_pos"].numel()])
    dist.broadcast(header, src=src, group=group)
    n_vec = int(header[0])
    plan = plan_shards(n_vec, world)
    sizes = header[1:].reshape(world, 4).tolist()
    val_dtype = torch.int64 if value_bytes == 8 else torch.int32
    if rank == src:
        reqs = []
        for r, s in enumerate(shards):
            if r == src:
                continue
            for key in ("meta", "packed", "exc_val", "exc_pos"):
                if s[key].numel():
                    # as bytes: NCCL has no 16-bit integer type (exc_pos)
                    reqs.append(dist.isend(s[key].contiguous().view(torch.uint8).reshape(-1), dst=r, group=group))
        for q in reqs:
            q.wait()
        mine = shards[src]
    else:
        nm, npk, nev, nep = sizes[rank]
        mine = {
            "meta": torch.empty((nm, 32), dtype=torch.uint8, device=device),
            "packed": torch.empty(npk, dtype=torch.uint8, device=device),
            "exc_val": torch.empty(nev, dtype=val_dtype, device=device),
            "exc_pos": torch.empty(nep, dtype=torch.int16, device=device),
        }
        for key in ("meta", "packed", "exc_val", "exc_pos"):
            if mine[key].numel():
                dist.recv(mine[key].view(torch.uint8).reshape(-1), src=src, group=group)
    return mine, plan[rank]


def column_tensors(dev_col):
    """DeviceColumn -> dict of tensors trimmed to the used sizes (what scatter_column / slice_column take)."""
    packed_bytes, n_exc = dev_col.read_totals()
    return {"meta": dev_col.meta, "packed": dev_col.packed[:packed_bytes], "exc_val": dev_col.exc_val[:n_exc], "exc_pos": dev_col.exc_pos[:n_exc]}


def host_column_tensors(h):
    """HostColumn -> dict of CPU tensors."""
    signed = np.int64 if h.value_bytes == 8 else np.int32
    return {
        "meta": torch.from_numpy(h.meta.view(np.uint8).reshape(-1, 32).copy()),
        "packed": torch.from_numpy(np.ascontiguousarray(h.packed[: h.packed_bytes]).copy()),
        "exc_val": torch.from_numpy(h.exc_val[: h.n_exceptions].view(signed).copy()),
        "exc_pos": torch.from_numpy(h.exc_pos[: h.n_exceptions].view(np.int16).copy()),
    }


def tensors_to_host_column(t, value_bytes):
    """dict of tensors (any device) -> HostColumn."""
    n_vec = t["meta"].shape[0]
    packed = t["packed"].cpu().numpy()
    h = _abi.HostColumn(n_vec, value_bytes, max(packed.shape[0], 128), max(t["exc_pos"].numel(), 1))
    h.meta[:] = t["meta"].cpu().numpy().view(_abi.VEC_META_DTYPE).reshape(-1)
    h.packed[: packed.shape[0]] = packed
    n_exc = t["exc_pos"].numel()
    h.exc_val[:n_exc] = t["exc_val"].cpu().numpy().view(h.exc_val.dtype)
    h.exc_pos[:n_exc] = t["exc_pos"].cpu().numpy().view(np.uint16)
    h.totals[:] = [packed.shape[0], n_exc, 0, 0]
    return h


def tensors_to_device_column(t, value_bytes, device):
    """dict of CUDA tensors -> DeviceColumn sharing no storage with the input (decode-ready)."""
    from .codec import DeviceColumn

    n_vec = t["meta"].shape[0]
    col = DeviceColumn(n_vec, value_bytes, device, max(t["packed"].numel(), 128), max(t["exc_pos"].numel(), 1))
    col.meta.copy_(t["meta"])
    col.packed[: t["packed"].numel()].copy_(t["packed"])
    col.exc_val[: t["exc_val"].numel()].copy_(t["exc_val"])
    col.exc_pos[: t["exc_pos"].numel()].copy_(t["exc_pos"])
    col.totals.copy_(torch.tensor([t["packed"].numel(), t["exc_pos"].numel(), 0, 0], dtype=torch.int64))
    return col
