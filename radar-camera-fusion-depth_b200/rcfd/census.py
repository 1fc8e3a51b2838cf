"""
Per-call timing of one step, measured live with CUDA events (measurement utility of bench.py / tools).

Every kernel of the hot path is launched through ``rcfd._lib.call`` (one C-ABI call = one kernel, plus a memset for a
few).  ``record()`` captures the (entry point, arguments) list of one eagerly executed step; ``replay()`` re-issues
each call on the current stream between two CUDA events, L2 flushed before each, so every kernel is timed alone and
warm-clock / cold-cache like the ncu launch list, but without a profiler attached.  Convolution calls are labelled
with the kernel the dispatcher picked (``rcfd_last_kernel``) and carry their algorithmic FLOPs / bytes, which is what
``roofline.achieved`` is computed from.
"""
import ctypes

import torch

from . import _lib, ops


class record(object):
    """with census.record() as calls: <run one eager step>  ->  calls = [(name, args, kernel label)].
    Tensors allocated by rcfd.ops during the step are kept alive until ``release()``."""

    def __enter__(self):
        self.hold = ops.hold_allocations()
        self.hold.__enter__()
        _lib.census = []
        self.calls = _lib.census
        return self

    def __exit__(self, *exc):
        _lib.census = None
        self.kept = list(ops._HOLD) if ops._HOLD is not None else []
        self.hold.__exit__(*exc)
        return False

    def release(self):
        self.kept = []
        self.calls = []


def conv_work(name, args):
    """(algorithmic FLOPs, executed FLOPs, algorithmic bytes) of a rcfd_conv2d_fwd / rcfd_conv2d_wgrad call.
    Algorithmic = the convolution the reference computes (2 x MACs on non-structural-zero inputs: an up-sampled 3x3
    conv counts its 9 taps, a stride-2 dgrad its real taps); executed = what the kernel multiplies (4/9 of the MACs
    on the sub-pixel path); bytes = every source / destination tensor once + the weights."""
    d = args[0]._obj
    cin = d.c0 + d.c1
    macs = float(d.n) * d.ho * d.wo * d.cout * d.kh * d.kw * cin / float(d.in_dilation * d.in_dilation)
    executed = macs * (4.0 / 9.0 if (d.weight_up2x and name == 'rcfd_conv2d_fwd') else 1.0)
    es = 2 if d.dtype == _lib.BF16 else 4
    src = float(d.n) * (d.h0 * d.w0 * d.c0 + d.hin * d.win * d.c1) * es
    dst = float(d.n) * d.ho * d.wo * d.cout * (4 if (d.dst_f32 and name == 'rcfd_conv2d_fwd') else es)
    wbytes = float(d.cout) * d.kh * d.kw * cin * (es if name == 'rcfd_conv2d_fwd' else 4)
    return 2.0 * macs, 2.0 * executed, src + dst + wbytes


def replay(calls, reps=1, flush_bytes=256 << 20):
    """Time every recorded call alone.  Returns a list of dicts (name, kernel, ms, flops, executed_flops, bytes)."""
    lib = _lib.load()
    dev = torch.cuda.current_device()
    flush = torch.empty(flush_bytes, dtype=torch.uint8, device='cuda:%d' % dev)
    out = []
    for name, args, kernel in calls:
        fn = getattr(lib, name)
        best = None
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            rc = fn(*args)
            b.record()
            if rc != 0:
                raise _lib.RcfdError('%s failed in replay: %s' % (name, lib.rcfd_last_error().decode()))
            b.synchronize()
            ms = a.elapsed_time(b)
            best = ms if best is None else min(best, ms)
        row = {'name': name, 'kernel': kernel, 'ms': best, 'flops': 0.0, 'executed_flops': 0.0, 'bytes': 0.0}
        if name in ('rcfd_conv2d_fwd', 'rcfd_conv2d_wgrad'):
            row['flops'], row['executed_flops'], row['bytes'] = conv_work(name, args)
        out.append(row)
    return out


def by_kernel(rows):
    """Aggregate replay rows per kernel label, sorted by total time (descending)."""
    agg = {}
    for r in rows:
        a = agg.setdefault(r['kernel'], {'kernel': r['kernel'], 'launches': 0, 'ms': 0.0, 'flops': 0.0,
                                         'executed_flops': 0.0, 'bytes': 0.0})
        a['launches'] += 1
        for k in ('ms', 'flops', 'executed_flops', 'bytes'):
            a[k] += r[k]
    total = sum(a['ms'] for a in agg.values()) or 1.0
    res = sorted(agg.values(), key=lambda a: -a['ms'])
    for a in res:
        a['share'] = a['ms'] / total
    return res


_CATEGORIES = (('conv fwd/dgrad', ('conv_',)), ('wgrad', ('wgrad_',)),
               ('batch norm', ('rcfd_bn_',)), ('pack/unpack', ('rcfd_pack_', 'rcfd_unpack_')))


def by_category(rows):
    total = sum(r['ms'] for r in rows) or 1.0
    out = {}
    for r in rows:
        cat = 'other'
        for label, prefixes in _CATEGORIES:
            if r['kernel'].startswith(prefixes):
                cat = label
                break
        out[cat] = out.get(cat, 0.0) + r['ms']
    return {k: {'ms': v, 'share': v / total} for k, v in sorted(out.items(), key=lambda kv: -kv[1])}
