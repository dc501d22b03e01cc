#!/usr/bin/env python
"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck): every kernel, small sizes."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alp_b200  # noqa: E402

dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
cols = [
    alp_b200.generate(3 * 102400, 2, dev),
    alp_b200.generate(2 * 102400 + 7 * 1024, 3, dev),
    alp_b200.generate(2 * 102400, 4, dev),
    torch.from_numpy(np.full(2048, np.nan)).to(dev),
    torch.from_numpy(rng.integers(0, 1 << 63, size=1024 * 9, dtype=np.uint64).view(np.float64)).to(dev),
    torch.from_numpy(rng.integers(0, 1 << 32, size=1024 * 9, dtype=np.uint64).astype(np.uint32).view(np.float32)).to(dev),
    torch.zeros(4096, dtype=torch.float32, device=dev),
]
# integers of every FFOR width (one vector each): the in-place packers, the wide direct path, the SUM fast / slow paths
cols.append(torch.from_numpy(np.concatenate([rng.integers(0, 1 << w, size=1024, dtype=np.uint64).astype(np.float64) for w in range(0, 53)])).to(dev))
cols.append(torch.from_numpy(np.concatenate([rng.integers(0, 1 << w, size=1024, dtype=np.uint64).astype(np.float32) for w in range(0, 24)])).to(dev))
# exception-heavy vectors (more than the encoder keeps in registers): the position-list path
kk = torch.from_numpy(rng.integers(0, 1000000, size=1024 * 12).astype(np.float64)).to(dev)
cols.append(kk / torch.tensor([1.0, 10.0, 100.0, 1000.0], dtype=torch.float64, device=dev)[torch.arange(1024 * 12, device=dev) % 4])
for x in cols:
    for ordered in (True, False):
        col = alp_b200.encode(x, ordered=ordered)
        col.read_totals()
        y = alp_b200.decode(col)
        ib = torch.int64 if x.element_size() == 8 else torch.int32
        assert torch.equal(x.view(ib), y.view(ib))
        s = alp_b200.decode_sum(col)
        s = alp_b200.decode_sum(col, flags=alp_b200.SUM_DECIMAL)
        alp_b200.minmax_result(alp_b200.decode_minmax(col))
        alp_b200.decode_filter(col, "<=", 10.0)
        torch.cuda.synchronize()
from alp_b200 import primitives as gpu  # noqa: E402

v = (rng.integers(0, 100000, 1024) / 100.0)
st = gpu.init(v)
r = gpu.encode(v, st)
bw, base = gpu.analyze_ffor(r["enc"])
p = gpu.ffor(r["enc"].view(np.uint64), bw, int(base))
d = gpu.patch(gpu.falp(p, bw, int(base), r["f"], r["e"]), r["exc"], r["pos"])
assert d.tobytes() == v.tobytes()
b8 = rng.integers(0, 32, 1024).astype(np.uint8)
assert gpu.unffor(gpu.ffor(b8, 5, 0), 5, 0, np.uint8).tobytes() == b8.tobytes()
codec = alp_b200.HostCodec(300, 8)
h = codec.compress(np.ascontiguousarray(cols[0].cpu().numpy()[:200000]))
assert codec.decompress(h).tobytes() == cols[0].cpu().numpy()[:200000].tobytes()
codec.sum(h)
codec.close()
print("sanitize smoke OK")
