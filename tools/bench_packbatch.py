"""Times the two rcfd_pack_batch launches of a FusionNet training step in isolation (CUDA events; L2 flushed or warm),
whole table and per item kind.    python tools/bench_packbatch.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'radar-camera-fusion-depth_b200'))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import fusionnet_model  # noqa: E402
from rcfd import ops, optim, synth  # noqa: E402

dev = torch.device('cuda:0')
m = fusionnet_model.FusionNetModel(device=dev, **synth.CANONICAL_FUSIONNET)
m.set_precision('bf16')
m.train()
opt = optim.FusedAdam(m.parameters(), lr=1e-3)
data = [t.to(dev) for t in bench.synthetic_batch(2, 0)]
for _ in range(2):
    d = m.forward(data[0], data[1])
    loss, _ = m.compute_loss(data[0], d, data[2], data[3], 'l1', 0.0, -1, None, 2.0)
    loss.backward()
torch.cuda.synchronize()
pack = m._cache[('pack_batch', torch.bfloat16)]['table']
unpack = m._cache[('unpack_state', torch.bfloat16)]['batch']['table']
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(table, cold, reps=10):
    tot = 0.0
    for _ in range(reps):
        if cold:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        table.run()
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / reps * 1e3


def subset(table, kinds):
    t = ops.PackBatch()
    t.keep = table.keep
    for it in table.rows:
        if it.kind in kinds:
            c = type(it)()
            for f, _ in it._fields_:
                setattr(c, f, getattr(it, f))
            c.block0 = t.total_blocks
            t.rows.append(c)
            t.total_blocks += c.nblocks
    return t.finalize(dev) if t.rows else None


names = {0: 'fwd', 1: 'dgrad', 2: 'up2x', 3: 'stem', 4: 'unpack', 5: 'unpack_stem', 6: 'copy', 7: 'dgrad_s2', 8: 'upconv_dgrad'}
for label, table in (('pack', pack), ('unpack', unpack)):
    elems = sum(it.total for it in table.rows)
    print('%s: %d items, %d blocks, %.1f M elements: cold %.1f us, warm %.1f us' %
          (label, len(table.rows), table.total_blocks, elems / 1e6, timed(table, True), timed(table, False)))
    for k in sorted(set(it.kind for it in table.rows)):
        sub = subset(table, (k,))
        print('   %-12s %4d items %6d blocks %6.1f M elements: cold %.1f us, warm %.1f us' %
              (names.get(k, str(k)), len(sub.rows), sub.total_blocks, sum(it.total for it in sub.rows) / 1e6, timed(sub, True), timed(sub, False)))
