"""
Executor of the FusionNet / RadarNet graphs on librcfd_b200.so.

The module trees in networks.py / net_utils.py hold parameters in the reference's layout;
this file walks them and issues C-ABI calls (rcfd.ops).  Activations are NHWC in the
compute dtype (float32 "parity" mode or bfloat16 "fast" mode), accumulation is fp32.

What the reference does as separate library calls is fused here:
  * eval mode:  conv + folded BN + activation (+ residual add + activation, + depth head)
                = ONE kernel per conv (reference src/net_utils.py:84-91, :323);
  * train mode: conv epilogue accumulates the BatchNorm batch statistics, then one
                normalise+activation(+residual) pass;
  * nearest up-sampling and channel concat are address maps of the consuming conv
    (reference src/net_utils.py:196, :565), never materialised;
  * the two 1x1 fusion convs of a level run as one stacked GEMM followed by one
    gate kernel (reference src/networks.py:864-866).

Training keeps a tape of closures (reverse-mode), so ``loss.backward()`` on the tensor
returned by FusionNetModel.forward runs dgrad / wgrad / BN-backward kernels and writes
parameter gradients (float32 OIHW, the reference's layout).
"""
import torch

from . import ops
from .ops import ACT_NONE, ACT_LEAKY, ACT_SIGMOID, ACT_DEPTH_HEAD

_ACT = {'linear': ACT_NONE, 'leaky_relu': ACT_LEAKY, 'sigmoid': ACT_SIGMOID}
CPAD = 16    # RGB image / depth+response are stored with 16 channels (TMA boxes and UMMA K need 32 B rows)
# the per-tap engine needs >= 64 dy channels per tap to beat `dgrad at the up-sampled resolution + 2x2 sum` (32-channel taps
# are 12 KB loads behind a full barrier round: measured no gain on deconv0)
import os as _os
UPCONV_DGRAD_MIN_C = int(_os.environ.get('RCFD_UPCONV_DGRAD_MIN_C', '64'))
CPAD_DY = 16  # d(logit) of the 1-channel head: 16 channels (32-byte rows) so its dgrad / wgrad stream through the row engines


_PARAM_EPOCH = [0]


def note_params_changed():
    """Invalidate every packed-weight / folded-BN cache entry (parameters were written outside torch)."""
    _PARAM_EPOCH[0] += 1


# Stream roles of the multi-stream schedule (one CUDA stream each; index 0 is the caller's current stream):
#   MAIN   image branch of the encoder, decoder, loss          DEPTH  depth branch of the encoder
#   FUSE   gated fusion levels (need both branches)            WG_*   weight gradients of the MAIN / DEPTH chains
#   PACK   weight packing of the whole step, issued up front (training)
# Most FusionNet layers below 1/4 resolution occupy a fraction of the 148 SMs and are latency bound; the
# chains are independent between fusion points, so running them side by side (and capturing them as
# parallel branches of the step's CUDA graph) overlaps those latencies.
MAIN, DEPTH, FUSE, WG_MAIN, WG_DEPTH, PACK, WG_X0, WG_X1 = 0, 1, 2, 3, 4, 5, 6, 7
_WG_OF = {MAIN: WG_MAIN, DEPTH: WG_DEPTH}
# weight gradients go round-robin over four streams: the last ones of a step (stems, level-1 / level-2 layers: 80-110 us each)
# have nothing left to hide behind, so they should at least run side by side instead of queueing on two streams
_WG_POOL = (WG_MAIN, WG_DEPTH, WG_X0, WG_X1)
_SIDE_STREAMS = {}


def side_streams(device):
    """The five side streams of a device (created once, outside any graph capture)."""
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    if key not in _SIDE_STREAMS:
        # DEPTH / FUSE carry the step's dependency chains: high priority, so that their CTAs are scheduled before those
        # of the weight-gradient / packing kernels (nothing waits for those before the optimiser), which otherwise fill
        # all 148 SMs in front of a critical-path kernel.  RCFD_STREAM_PRIORITY=0 switches it off.
        import os
        hi = -1 if os.environ.get('RCFD_STREAM_PRIORITY', '1') != '0' else 0
        _SIDE_STREAMS[key] = [torch.cuda.Stream(device=device, priority=hi if i < 2 else 0) for i in range(7)]
    return _SIDE_STREAMS[key]


_CAPTURE_STREAMS = {}


def capture_stream(device):
    """High-priority stream to capture the step graphs on (the MAIN role of a captured step)."""
    import os
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    if key not in _CAPTURE_STREAMS:
        hi = -1 if os.environ.get('RCFD_STREAM_PRIORITY', '1') != '0' else 0
        _CAPTURE_STREAMS[key] = torch.cuda.Stream(device=device, priority=hi)
    return _CAPTURE_STREAMS[key]


class Tape(object):
    """Reverse-mode tape: forward ops append closures; grads are keyed by tensor identity.
    With ``streams`` (multi-stream schedule) every step replays on the stream role it was recorded on and
    gradients that cross roles carry the event of their last writer."""

    def __init__(self, streams=None):
        self.steps = []              # (closure, stream role)
        self.grads = {}              # id(tensor) -> [grad, role of last writer, event of last write]
        self.keep = []
        self.param_grads = []        # (parameter, grad tensor float32 in the parameter's layout)
        self.streams = streams       # None = single stream
        self.sid = MAIN              # role being recorded (forward) / replayed (backward)
        self.finalizers = []         # run on the caller's stream once the backward streams have joined
        self.split_at = None         # index into steps: backward(part=0) replays steps[split_at:], part=1 the rest
        self.mid_finalizers = []     # run at the end of part 0

    def add_step(self, fn):
        self.steps.append((fn, self.sid))

    def _mark(self):
        if self.streams is None:
            return None
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        return ev

    def _wait(self, entry):
        if self.streams is not None and entry[2] is not None and entry[1] != self.sid:
            torch.cuda.current_stream().wait_event(entry[2])

    def set_grad(self, t, g):
        """Seed (the loss gradient): produced on the caller's stream before backward() starts."""
        self.grads[id(t)] = [g, MAIN, None]
        self.keep.append(t)

    def add_grad(self, t, g):
        """Publish a gradient contribution.  Call it AFTER the last read of ``g`` on the current stream:
        a later contribution from another role accumulates into ``g`` in place."""
        k = id(t)
        e = self.grads.get(k)
        if e is not None:
            self._wait(e)
            ops.add_(e[0], g)
            e[1], e[2] = self.sid, self._mark()
        else:
            self.grads[k] = [g, self.sid, self._mark()]
            self.keep.append(t)

    def grad_of(self, t):
        e = self.grads.pop(id(t), None)
        if e is None:
            return None
        self._wait(e)
        return e[0]

    def side(self, fn):
        """Run ``fn`` (a weight gradient: nothing downstream of it but the optimiser) on the weight-gradient
        stream of the current role, ordered after everything enqueued so far on the current stream."""
        if self.streams is None:
            fn()
            return
        self.wg_rr = getattr(self, 'wg_rr', 0) + 1
        ws = self.streams[_WG_POOL[self.wg_rr % len(_WG_POOL)]]
        ws.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(ws):
            fn()

    def mark_split(self):
        """Forward-time marker: everything recorded AFTER this call is replayed by backward(part=0), the rest by part=1
        (data parallel: the gradients of the layers behind the marker are all-reduced while part 1 runs)."""
        self.split_at = len(self.steps)

    def backward(self, part=None):
        """part=None: the whole tape.  part=0 / part=1: the two halves around mark_split(); every stream is joined at the
        end of part 0, so part 1 may be captured into a second CUDA graph."""
        if part is not None and self.split_at is None:
            raise RuntimeError('Tape.backward(part): no split marker was recorded')
        steps = self.steps if part is None else (self.steps[self.split_at:] if part == 0 else self.steps[:self.split_at])
        last = part is None or part == 1
        if self.streams is None:
            for fn, _ in reversed(steps):
                fn()
            for fn in (self.finalizers if last else self.mid_finalizers):
                fn()
        else:
            main = torch.cuda.current_stream()
            self.streams[MAIN] = main
            with ops.hold_allocations():
                for s in self.streams[1:]:
                    s.wait_stream(main)
                for fn, sid in reversed(steps):
                    self.sid = sid
                    if sid == MAIN:
                        fn()
                    else:
                        with torch.cuda.stream(self.streams[sid]):
                            fn()
                self.sid = MAIN
                for s in self.streams[1:]:
                    main.wait_stream(s)
                for fn in (self.finalizers if last else self.mid_finalizers):
                    fn()
        if not last:
            for e in self.grads.values():          # everything is joined: later waits must not reference this part's events
                e[2] = None
            return
        self.finalizers = []
        self.mid_finalizers = []
        self.steps = []
        self.grads = {}
        self.keep = []


class Context(object):
    """Per-forward state: dtype, mode, tape, BatchNorm scratch pool, packed-weight cache."""

    def __init__(self, dtype, training, device, cache=None, record=False, engine=ops.ENGINE_AUTO, multistream=False,
                 x3=False, external_pack=False):
        self.dtype = dtype
        # tensor-core parity modes: fp32 storage (dtype float32), every conv / wgrad = 3 or 6 bf16 tcgen05 passes over
        # 2- or 3-part bf16 operand splits (rcfd/x3.py); x3 = number of parts (0: off), packed weights are tuples of parts
        self.x3 = int(x3)
        assert not self.x3 or dtype == torch.float32
        self.training = training
        self.device = device
        # multi-stream schedule: [caller's stream, DEPTH, FUSE, WG_MAIN, WG_DEPTH, PACK]
        multistream = multistream and torch.device(device).type == 'cuda'
        self.streams = [torch.cuda.current_stream()] + side_streams(device) if multistream else None
        self.sid = MAIN
        self.tape = Tape(self.streams) if record else None
        self.cache = cache if cache is not None else {}
        self.engine = engine
        self._pool_stats = None
        self._pool_aff = None
        self._off_stats = 0
        self._off_aff = 0
        self.bn_counters = []
        self.taps = None
        self.prepacked = {}
        # external_pack: the caller runs the batched pack launch itself before this step (FusionNetModel.train_step_graphed
        # issues it right after the optimiser step, where it overlaps the launch latency of the next graph replay)
        self.external_pack = external_pack
        # data parallel with an overlapped gradient all-reduce: address in the flat gradient buffer from which on the
        # gradients belong to the layers behind the tape's split marker (encoder levels >= 5, decoder); None = no split
        self.grad_split = None
        # pack plan / batched (un)packing: every training context on the device (also the single-stream schedule, so that
        # profiles taken on one stream show the kernels the graphed step runs)
        on_device = torch.device(device).type == 'cuda'
        self.plan = self.cache.setdefault(('pack_plan', dtype, self.x3), {}) if (on_device and training) else None
        # batched weight (un)packing (training, multi-stream, plain dtypes): ONE rcfd_pack_batch launch per step packs every
        # weight the step uses into persistent buffers, another one unpacks every weight gradient at the end of backward
        self.unpack = None
        if self.plan is not None and not self.x3 and self.tape is not None:
            self.unpack = self.cache.setdefault(('unpack_state', dtype), {'bufs': {}, 'items': {}, 'batch': None, 'dirty': False})
            _validate_unpack(self.unpack)
            self.tape.finalizers.append(lambda: _finish_unpack(self.unpack, self.device, self.grad_split, 1))
            self.tape.mid_finalizers.append(lambda: _finish_unpack(self.unpack, self.device, self.grad_split, 0))
            # the pack table of the NEXT step is built here, at the end of the step that recorded the plan (eagerly: a
            # capture that follows, FusionNetModel.train_step_graphed, then already replays the batched form)
            self.tape.finalizers.append(lambda: _pack_batch_for(self.cache, self.plan, self.dtype, self.device))
        if self.streams is not None and training:
            self.stats(0)            # the zeroed statistics pool must exist before the streams fork

    # -- stream roles (no-ops on the single-stream schedule)
    def fork(self):
        """Side streams start after everything enqueued so far on the caller's stream."""
        if self.streams is not None:
            for s in self.streams[1:3]:
                s.wait_stream(self.streams[MAIN])

    def join(self):
        if self.streams is not None:
            for s in self.streams[1:3] + ([self.streams[PACK]] if self.prepacked or self.plan else []):
                self.streams[MAIN].wait_stream(s)

    def wait(self, role, on):
        """Stream ``role`` waits for what has been enqueued so far on streams ``on``."""
        if self.streams is not None:
            for o in on:
                if o != role:
                    self.streams[role].wait_stream(self.streams[o])

    def on(self, role):
        return _Role(self, role)

    # -- scratch: stats are zero-initialised doubles, affine params are floats
    def stats(self, c):
        if self._pool_stats is None or self._off_stats + 2 * c > self._pool_stats.numel():
            if self._pool_stats is not None and self.streams is not None:
                raise RuntimeError('BatchNorm statistics pool exhausted under the multi-stream schedule')
            self._pool_stats = torch.zeros(max(65536, 2 * c), device=self.device, dtype=torch.float64)
            self._off_stats = 0
        s = self._pool_stats[self._off_stats:self._off_stats + 2 * c]
        self._off_stats += 2 * c
        return s[:c], s[c:]

    def zeroed_sums(self, c):
        """Zeroed float64 [2c] scratch for a BatchNorm backward reduction (multi-stream training: a slice of the step's
        one zeroed pool, so the backward chains carry no memset nodes), else None (the kernel call zeroes its own)."""
        if self.streams is None or not self.training:
            return None
        a, b = self.stats(c)
        return self._pool_stats[a.storage_offset():a.storage_offset() + 2 * c]

    def aff(self, c, k=4):
        n = k * c
        if self._pool_aff is None or self._off_aff + n > self._pool_aff.numel():
            self._pool_aff = torch.empty(max(131072, n), device=self.device, dtype=torch.float32)
            self._off_aff = 0
        s = self._pool_aff[self._off_aff:self._off_aff + n]
        self._off_aff += n
        return [s[i * c:(i + 1) * c] for i in range(k)]

    # -- weights: packed per step in training (they change every step), cached by version in eval.
    #    Optimisers that write parameters through raw pointers (rcfd.optim.FusedAdam) do not move torch's
    #    version counters: they call note_params_changed(), which is part of the cache key.
    #    Training under the multi-stream schedule: the first step records every pack call (key -> closure) in
    #    the model's pack plan; later steps replay the plan up front on the PACK stream (prepack), so the ~150
    #    tiny pack kernels leave the critical chains and each consumer just waits for its event.
    def packed(self, key, params, fn, spec=None):
        """spec: callable returning the ops.spec_pack_* description of what fn does (lets the step's packs run as one
        batched launch, see prepack)."""
        if self.x3 and not key[0].startswith('bn'):
            fn = _split_after(fn, self.x3)       # pack in fp32 (self.dtype), then split the packed tensor into bf16 parts
            spec = None
        if self.training:
            hit = self.prepacked.pop(key, None)
            if hit is not None:
                val, ev = hit
                if ev is not None:
                    torch.cuda.current_stream().wait_event(ev)
                return val
            if self.plan is not None:
                self.plan[key] = (fn, spec)
            return fn()
        ver = tuple(p._version for p in params) + (self.dtype, self.x3, _PARAM_EPOCH[0])
        hit = self.cache.get(key)
        if hit is not None and hit[0] == ver:
            return hit[1]
        val = fn()
        self.cache[key] = (ver, val)
        return val

    def prepack(self):
        """Issue every pack call recorded by the previous step on the PACK stream (training, multi-stream)."""
        if self.plan is None or not self.plan:
            return
        if self.streams is not None:
            pk, main = self.streams[PACK], self.streams[MAIN]
            pk.wait_stream(main)
        else:
            pk = torch.cuda.current_stream()
        with torch.cuda.stream(pk):
            batch = _pack_batch_for(self.cache, self.plan, self.dtype, self.device)
            if batch is not None:
                ev = None
                if not self.external_pack:
                    batch['table'].run()
                    ev = torch.cuda.Event()
                    ev.record(pk)
                for key, out in batch['outputs'].items():
                    self.prepacked[key] = (out, ev)
            for key, (fn, spec) in self.plan.items():
                if batch is not None and key in batch['outputs']:
                    continue
                val = fn()
                ev = torch.cuda.Event()
                ev.record(pk)
                self.prepacked[key] = (val, ev)

    def weight(self, mod, pad_to=None):
        """Packed [cout][taps][cin_pad] weights; pad_to = channel count of the (zero-padded) input."""
        w, dtype = mod.conv.weight, self.dtype          # the closures outlive this context (pack plan): no self in them
        return self.packed(('w', id(mod), pad_to), [w], lambda: ops.pack_weight(w.detach(), dtype, pad_to=pad_to),
                           spec=lambda: ops.spec_pack_weight(w.detach(), dtype, pad_to=pad_to))

    def weight_stem_s2d(self, mod):
        """7x7/s2 stem weights rearranged for the space-to-depth (4x4/s1) formulation."""
        w, dtype = mod.conv.weight, self.dtype
        return self.packed(('ws2d', id(mod)), [w], lambda: ops.pack_stem_s2d_weight(w.detach(), dtype, CPAD),
                           spec=lambda: ops.spec_pack_stem_s2d_weight(w.detach(), dtype, CPAD))

    def weight_up2x(self, mod):
        """Sub-pixel phase weights for a 3x3 conv behind an exact 2x nearest up-sampling (bf16 fast path)."""
        w, dtype = mod.conv.weight, self.dtype
        return self.packed(('wup', id(mod)), [w], lambda: ops.pack_upconv2x_weight(w.detach(), dtype),
                           spec=lambda: ops.spec_pack_upconv2x_weight(w.detach(), dtype))

    def weight_dgrad(self, mod, off, cnt, pad_to):
        """Packed [cin_cnt][taps][cout_pad] weights of the data-gradient convolution (one concat source)."""
        w, dtype = mod.conv.weight, self.dtype
        return self.packed(('wd', id(mod), off, cnt, pad_to), [w],
                           lambda: ops.pack_weight(w.detach(), dtype, cin_off=off, cin_cnt=cnt, dgrad=True, pad_to=pad_to),
                           spec=lambda: ops.spec_pack_weight(w.detach(), dtype, cin_off=off, cin_cnt=cnt, dgrad=True,
                                                             pad_to=pad_to))

    def weight_dgrad_s2(self, mod, off, cnt, pad_to):
        """Phase weights of the zero-insertion-free data gradient of a 3x3 / stride-2 conv (bf16 fast path)."""
        w, dtype = mod.conv.weight, self.dtype
        return self.packed(('wds2', id(mod), off, cnt, pad_to), [w],
                           lambda: ops.pack_dgrad_s2_weight(w.detach(), dtype, cin_off=off, cin_cnt=cnt, pad_to=pad_to),
                           spec=lambda: ops.spec_pack_dgrad_s2_weight(w.detach(), dtype, cin_off=off, cin_cnt=cnt,
                                                                      pad_to=pad_to))

    def weight_upconv_dgrad(self, mod, off, cnt, pad_to):
        """4x4 / stride-2 weights of the data gradient of an exact-2x up-conv w.r.t. its low-res source (bf16 fast path)."""
        w, dtype = mod.conv.weight, self.dtype
        return self.packed(('wupd', id(mod), off, cnt, pad_to), [w],
                           lambda: ops.pack_upconv2x_dgrad_weight(w.detach(), dtype, cin_off=off, cin_cnt=cnt, pad_to=pad_to),
                           spec=lambda: ops.spec_pack_upconv2x_dgrad_weight(w.detach(), dtype, cin_off=off, cin_cnt=cnt,
                                                                            pad_to=pad_to))

    def folded_bn(self, mod):
        bn = mod.batch_norm
        # num_batches_tracked: the training kernels update the running statistics through raw pointers (no version bump)
        ps = [bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.num_batches_tracked]

        def make():
            scale = torch.empty(bn.num_features, device=self.device, dtype=torch.float32)
            shift = torch.empty_like(scale)
            ops.bn_fold(bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var, scale, shift)
            return scale, shift
        return self.packed(('bn', id(mod)), ps, make)


def _split_after(fn, parts):
    return lambda: ops.split_bf16(fn(), parts)


def _pack_batch_for(cache, plan, dtype, device):
    """The batched form of a pack plan: persistent destination tensors + one device table (rcfd_pack_batch), built once
    per model / dtype outside graph capture and reused while the plan and the parameters' storage stay the same."""
    key = ('pack_batch', dtype)
    batch = cache.get(key)
    if batch is not None:
        # the plan grew, or a parameter's storage moved (model.to(), rcfd.optim.FusedAdam re-pointing .data into its flat
        # buffer): the specs are re-evaluated (they take fresh aliases of the parameters) and compared with the table
        now = [it['src'].data_ptr() for _, (fn, spec) in plan.items() if spec is not None for it in spec()[3]] \
            if batch['n_plan'] == len(plan) else None
        if now != batch['src_ptrs']:
            batch = None
    if batch is None:
        if torch.cuda.is_current_stream_capturing():
            return None                  # table upload and allocations belong outside a capture: per-item path this time
        table, outputs, srcs = ops.PackBatch(), {}, []
        src_ptrs = []
        for k, (fn, spec) in plan.items():
            if spec is None:
                continue
            shape, dt_, zero, items = spec()
            out = (torch.zeros if zero else torch.empty)(shape, device=device, dtype=dt_)
            for it in items:
                it = dict(it)
                src = it.pop('src')
                table.add(it.pop('kind'), src, out, **it)
                srcs.append((src, src.data_ptr()))
                src_ptrs.append(src.data_ptr())
            outputs[k] = out
        if not outputs:
            return None
        batch = {'table': table.finalize(device), 'outputs': outputs, 'srcs': srcs, 'src_ptrs': src_ptrs, 'n_plan': len(plan)}
        cache[key] = batch
    return batch


def _validate_unpack(st):
    """Drop the batched-unpack table when a gradient destination moved (new optimiser / flat buffer)."""
    if st['batch'] is not None and any(p.grad is None or p.grad.data_ptr() != ptr for p, ptr in st['batch']['checks']):
        st['batch'] = None
        st['items'].clear()
        st['bufs'].clear()


def _wgrad_slot(ctx, key, shape, param, gw):
    """Persistent packed-gradient buffer for the batched unpack, or None when this gradient takes the per-layer path
    (no flat destination, parity modes).  Returns (buffer, deferred): deferred = the step's batch unpacks it."""
    st = ctx.unpack
    if st is None or gw is not param.grad or not getattr(param, '_rcfd_flat', False):
        return None, False
    buf = st['bufs'].get(key)
    if buf is None or tuple(buf.shape) != tuple(shape):
        buf = torch.empty(shape, device=ctx.device, dtype=torch.float32)
        st['bufs'][key] = buf
        st['dirty'] = True
    batch = st['batch']
    return buf, batch is not None and key in batch['keys']


def _finish_unpack(st, device, grad_split=None, part=1):
    """End of backward (streams joined): the one batched unpack; (re)build the table after a recording step.  With a
    gradient split (two-part backward) the table exists twice: destinations at / above the split address are unpacked at
    the end of part 0, the others at the end of part 1."""
    batch = st['batch']
    if batch is not None:
        if grad_split is None:
            batch['table'].run()
        else:
            tables = batch.get(('split', grad_split))
            if tables is None:
                if torch.cuda.is_current_stream_capturing():
                    raise RuntimeError('the split unpack tables must exist before a capture (run one eager step first)')
                tables = [ops.PackBatch(), ops.PackBatch()]
                for kind, src, src_off, dst, f in batch['rows']:
                    tables[0 if dst.data_ptr() >= grad_split else 1].add(kind, src, dst, src_off=src_off, **f)
                tables = [t.finalize(device) if t.rows else None for t in tables]
                batch[('split', grad_split)] = tables
            if tables[part] is not None:
                tables[part].run()
    if part == 1 and st['dirty'] and not torch.cuda.is_current_stream_capturing():
        table, checks, rows = ops.PackBatch(), [], []
        for key, items in st['items'].items():
            for kind, src, src_off, dst, param, f in items:
                table.add(kind, src, dst, src_off=src_off, **f)
                checks.append((param, dst.data_ptr()))
                rows.append((kind, src, src_off, dst, f))
        st['batch'] = {'table': table.finalize(device), 'keys': set(st['items']), 'checks': checks, 'rows': rows} if checks else None
        st['dirty'] = False


def _unpack_conv(ctx, key, dw, gw, param, src_off=0, rows=None):
    """Register (first time) the unpack of packed gradient ``dw`` (persistent) into ``gw`` for the batch."""
    st = ctx.unpack
    cout, cin, kh, kw = gw.shape
    cpad = dw.shape[2]
    f = dict(total=cout * min(cpad, cin) * kh * kw, cout=cout, cin=cin, taps=kh * kw, cin_off=0, cin_cnt=min(cpad, cin),
             cpad=cpad)
    st['items'].setdefault(key, []).append((ops.UNPACK_CONV, dw, src_off, gw, param, f))


class _Role(object):
    """``with ctx.on(role):`` -- issue (and record on the tape) the enclosed ops on that stream role."""

    def __init__(self, ctx, role):
        self.ctx, self.role = ctx, role

    def __enter__(self):
        ctx = self.ctx
        self.prev = ctx.sid
        self.cm = None
        if ctx.streams is not None and self.role != ctx.sid:
            ctx.sid = self.role
            if ctx.tape is not None:
                ctx.tape.sid = self.role
            self.cm = torch.cuda.stream(ctx.streams[self.role])
            self.cm.__enter__()
        return self

    def __exit__(self, *exc):
        ctx = self.ctx
        if self.cm is not None:
            self.cm.__exit__(*exc)
            ctx.sid = self.prev
            if ctx.tape is not None:
                ctx.tape.sid = self.prev
        return False


# ----------------------------------------------------------------------------- conv unit
def conv_unit(ctx, mod, x0, x1=None, in_size=None, residual=None, head=None, want_input_grad=True, stem_s2d=False):
    """net_utils.Conv2d: conv -> BN? -> act? (-> residual add -> leaky).  head=(min, min/max)
    turns the activation into the bounded depth head and stores float32."""
    k, stride, cout = mod.kernel_size, mod.stride, mod.out_channels
    act = _ACT[mod.act_kind]
    if stem_s2d:
        # 7x7 / stride-2 stem on a space-to-depth input == 4x4 / stride-1 / pad-2 conv (include/rcfd.h)
        assert k == 7 and stride == 2 and x1 is None and in_size is None and residual is None and head is None
        return _stem_s2d_unit(ctx, mod, x0, act)
    cin_data = x0.shape[3] + (x1.shape[3] if x1 is not None else 0)
    w = ctx.weight(mod, pad_to=cin_data if cin_data != mod.in_channels else None)   # 3/2-channel inputs live padded
    wup = None
    if (in_size is not None and x1 is None and k == 3 and stride == 1 and (ctx.dtype == torch.bfloat16 or ctx.x3)
            and int(in_size[0]) == 2 * x0.shape[1] and int(in_size[1]) == 2 * x0.shape[2] and x0.shape[3] % 16 == 0):
        wup = ctx.weight_up2x(mod)      # TMA engine: four 2x2 convs on the low-res source instead of a gather
    if head is not None:
        out = ops.conv2d(x0, w, cout, k, stride, x1=x1, in_size=in_size, act=ACT_DEPTH_HEAD, act_params=head,
                         out_f32=True, engine=ctx.engine)
        if ctx.tape is not None:
            _record_conv_backward(ctx, mod, x0, x1, in_size, out, None, want_input_grad,
                                  pre=lambda dd: ops.depth_head_bwd(dd, out, head[0], head[1], ctx.dtype,
                                                                    cpad=CPAD_DY if (ctx.dtype == torch.bfloat16 or ctx.x3) else 8))
        return out
    if not mod.use_batch_norm:
        out = ops.conv2d(x0, w, cout, k, stride, x1=x1, in_size=in_size, act=act, residual=residual, engine=ctx.engine)
        if ctx.tape is not None:
            assert act == ACT_NONE and residual is None
            _record_conv_backward(ctx, mod, x0, x1, in_size, out, None, want_input_grad)
        return out
    bn = mod.batch_norm
    if not ctx.training:
        scale, shift = ctx.folded_bn(mod)
        return ops.conv2d(x0, w, cout, k, stride, x1=x1, in_size=in_size, scale=scale, shift=shift, act=act,
                          residual=residual, engine=ctx.engine, weight_up2x=wup)
    # training: raw conv + batch statistics -> finalize -> normalise/activate(/residual)
    ssum, ssq = ctx.stats(cout)
    y = ops.conv2d(x0, w, cout, k, stride, x1=x1, in_size=in_size, stats=(ssum, ssq), engine=ctx.engine,
                   weight_up2x=wup)
    scale, shift, mean, invstd = ctx.aff(cout)
    z = ops.bn_train_act(y, ssum, ssq, bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var, scale, shift,
                         mean, invstd, act, residual=residual)
    ctx.bn_counters.append(bn.num_batches_tracked)
    if ctx.tape is not None:
        _record_conv_backward(ctx, mod, x0, x1, in_size, z, (y, scale, shift, mean, invstd, act, residual),
                              want_input_grad)
    return z


def _stem_s2d_unit(ctx, mod, x, act):
    cout = mod.out_channels
    hw = (x.shape[1], x.shape[2])
    w = ctx.weight_stem_s2d(mod)
    bn = mod.batch_norm
    if not ctx.training:
        scale, shift = ctx.folded_bn(mod)
        return ops.conv2d(x, w, cout, 4, 1, pad=2, out_size=hw, scale=scale, shift=shift, act=act, engine=ctx.engine)
    ssum, ssq = ctx.stats(cout)
    y = ops.conv2d(x, w, cout, 4, 1, pad=2, out_size=hw, stats=(ssum, ssq), engine=ctx.engine)
    scale, shift, mean, invstd = ctx.aff(cout)
    z = ops.bn_train_act(y, ssum, ssq, bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var, scale, shift,
                         mean, invstd, act)
    ctx.bn_counters.append(bn.num_batches_tracked)
    if ctx.tape is not None:
        tape = ctx.tape

        def bwd():
            dz = tape.grad_of(z)
            if dz is None:
                return
            dgamma, dbeta = _grad_dst(bn.weight), _grad_dst(bn.bias)
            dy = ops.bn_act_bwd(dz, y, scale, shift, mean, invstd, act, dgamma, dbeta, sums=ctx.zeroed_sums(cout))
            tape.param_grads.append((bn.weight, dgamma))
            tape.param_grads.append((bn.bias, dbeta))
            gw = _grad_dst(mod.conv.weight)

            def wgrad():
                ukey = ('dws2d', id(mod))
                buf, deferred = _wgrad_slot(ctx, ukey, (cout, 16, x.shape[3]), mod.conv.weight, gw)
                dw = ops.conv2d_wgrad(x, dy, 4, 1, pad=2, engine=ctx.engine, x3=ctx.x3, out=buf)      # [cout][16][CPAD]
                if not deferred:
                    ops.unpack_stem_s2d_wgrad(dw, gw)
                    if buf is not None and ukey not in ctx.unpack['items']:
                        c = gw.shape[1]
                        ctx.unpack['items'][ukey] = [(ops.UNPACK_STEM_S2D, dw, 0, gw, mod.conv.weight,
                                                      dict(total=cout * c * 49, cout=cout, cin=c, taps=16, cpad=dw.shape[2]))]
            tape.side(wgrad)
            tape.param_grads.append((mod.conv.weight, gw))
        tape.add_step(bwd)
    return z


def _grad_dst(param):
    """Where a parameter gradient is written: in place into the flat buffer view installed by
    rcfd.optim.FusedAdam, else a fresh tensor handed to FusionNetModel._deliver_grads."""
    if getattr(param, '_rcfd_flat', False) and param.grad is not None:
        return param.grad
    return torch.empty_like(param)


def _record_conv_backward(ctx, mod, x0, x1, in_size, z, bn_state, want_input_grad, pre=None):
    tape = ctx.tape
    k, stride = mod.kernel_size, mod.stride
    w_param = mod.conv.weight

    def bwd():
        dz = tape.grad_of(z)
        if dz is None:
            return
        if pre is not None:
            dz = pre(dz)
        if bn_state is not None:
            y, scale, shift, mean, invstd, act, residual = bn_state
            bn = mod.batch_norm
            dgamma = _grad_dst(bn.weight)
            dbeta = _grad_dst(bn.bias)
            if residual is not None:                    # through the post-add activation first (z = its output)
                dy, dz = ops.bn_act_bwd(dz, y, scale, shift, mean, invstd, act, dgamma, dbeta,
                                        sums=ctx.zeroed_sums(y.shape[3]), post_z=z)
            else:
                dy = ops.bn_act_bwd(dz, y, scale, shift, mean, invstd, act, dgamma, dbeta, sums=ctx.zeroed_sums(y.shape[3]))
            if residual is not None:
                tape.add_grad(residual, dz)             # published after its last read here (Tape.add_grad)
            tape.param_grads.append((bn.weight, dgamma))
            tape.param_grads.append((bn.bias, dbeta))
        else:
            dy = dz
        gw = _grad_dst(w_param)

        def wgrad():
            ukey = ('dw', id(mod))
            cin_tot = x0.shape[3] + (x1.shape[3] if x1 is not None else 0)
            buf, deferred = _wgrad_slot(ctx, ukey, (dy.shape[3], k * k, cin_tot), w_param, gw)   # d(logit) rows are padded
            dw = ops.conv2d_wgrad(x0, dy, k, stride, x1=x1, in_size=in_size, engine=ctx.engine, x3=ctx.x3, out=buf)
            if not deferred:
                ops.unpack_wgrad(dw, gw)
                if buf is not None and ukey not in ctx.unpack['items']:
                    _unpack_conv(ctx, ukey, dw, gw, w_param)
        tape.side(wgrad)
        tape.param_grads.append((w_param, gw))
        if not want_input_grad:
            return
        c0 = x0.shape[3]
        hin, win = (x0.shape[1], x0.shape[2]) if in_size is None else in_size
        pad_d = k - 1 - k // 2
        for (src, off, cnt) in ((x0, 0, c0),) + (((x1, c0, x1.shape[3]),) if x1 is not None else ()):
            if (src is x0 and (hin, win) == (2 * x0.shape[1], 2 * x0.shape[2]) and k == 3 and stride == 1
                    and ctx.dtype == torch.bfloat16 and not ctx.x3 and dy.shape[3] % UPCONV_DGRAD_MIN_C == 0 and cnt % 16 == 0
                    and ctx.engine == ops.ENGINE_AUTO):
                # gradient w.r.t. the low-res source of an exact-2x up-conv in one 4x4 / stride-2 conv over dy (no dgrad at
                # the up-sampled resolution, no 2x2 sum pass)
                w4 = ctx.weight_upconv_dgrad(mod, off, cnt, dy.shape[3])
                tape.add_grad(src, ops.conv2d(dy, w4, cnt, 4, 2, pad=1, engine=ctx.engine))
                continue
            if (stride == 2 and k == 3 and x1 is None and ctx.dtype == torch.bfloat16 and not ctx.x3
                    and dy.shape[3] % 16 == 0 and cnt % 16 == 0 and ctx.engine == ops.ENGINE_AUTO):
                # TMA engine: four 2x2 convs on the dy grid (one per destination parity) instead of 9 taps on the
                # zero-inserted grid through the gather engine
                wd = ctx.weight_dgrad_s2(mod, off, cnt, dy.shape[3])
                dsrc = ops.conv2d(dy, wd, cnt, k, 1, pad=pad_d, in_dilation=stride, out_size=(hin, win), engine=ctx.engine,
                                  weight_up2x=wd)
            else:
                wd = ctx.weight_dgrad(mod, off, cnt, dy.shape[3])
                dsrc = ops.conv2d(dy, wd, cnt, k, 1, pad=pad_d, in_dilation=stride, out_size=(hin, win), engine=ctx.engine)
            if src is x0 and (hin, win) != (x0.shape[1], x0.shape[2]):
                dsrc = ops.upsample_nearest_bwd(dsrc, (x0.shape[1], x0.shape[2]))
            tape.add_grad(src, dsrc)
    tape.add_step(bwd)


# ----------------------------------------------------------------------------- gated fusion level
def fusion_level(ctx, mod_w, mod_p, dep, img):
    """fused = sigmoid(BN(Ww . d)) * BN(Wp . d) + img   (reference src/networks.py:864-866)"""
    c = mod_w.out_channels
    ww, wp = mod_w.conv.weight, mod_p.conv.weight
    dtype = ctx.dtype
    wcat = ctx.packed(('wcat', id(mod_w)), [ww, wp],
                      lambda: ops.pack_weight(torch.cat([ww.detach(), wp.detach()], 0), dtype),
                      spec=lambda: ops.spec_pack_stacked_1x1([ww.detach(), wp.detach()], dtype))
    bw, bp = mod_w.batch_norm, mod_p.batch_norm
    if not ctx.training:
        def make():
            scale = torch.empty(2 * c, device=ctx.device, dtype=torch.float32)
            shift = torch.empty_like(scale)
            ops.bn_fold(bw.weight.detach(), bw.bias.detach(), bw.running_mean, bw.running_var, scale[:c], shift[:c])
            ops.bn_fold(bp.weight.detach(), bp.bias.detach(), bp.running_mean, bp.running_var, scale[c:], shift[c:])
            return scale, shift
        scale, shift = ctx.packed(('bncat', id(mod_w)),
                                  [bw.weight, bw.bias, bw.running_mean, bw.running_var,
                                   bp.weight, bp.bias, bp.running_mean, bp.running_var,
                                   bw.num_batches_tracked, bp.num_batches_tracked], make)
        y = ops.conv2d(dep, wcat, 2 * c, 1, 1, scale=scale, shift=shift, engine=ctx.engine)
        return ops.gate_fuse(y, None, None, img)
    ssum, ssq = ctx.stats(2 * c)
    y = ops.conv2d(dep, wcat, 2 * c, 1, 1, stats=(ssum, ssq), engine=ctx.engine)
    scale, shift, mean, invstd = ctx.aff(2 * c)
    count = y.numel() // (2 * c)
    for bn, lo in ((bw, 0), (bp, c)):
        ops.bn_finalize(ssum[lo:lo + c], ssq[lo:lo + c], bn.weight.detach(), bn.bias.detach(), bn.running_mean,
                        bn.running_var, scale[lo:lo + c], shift[lo:lo + c], mean[lo:lo + c], invstd[lo:lo + c], count)
        ctx.bn_counters.append(bn.num_batches_tracked)
    out = ops.gate_fuse(y, scale, shift, img)
    if ctx.tape is not None:
        tape = ctx.tape

        def bwd():
            dout = tape.grad_of(out)
            if dout is None:
                return
            dzy = ops.gate_fuse_bwd(dout, y, scale, shift)
            tape.add_grad(img, dout)
            # d(gamma) / d(beta) of the stacked pair land in one [2][2c] buffer; with flat gradient destinations the four
            # slices are copied by the step's batched unpack launch instead of four captured memcpys per level
            st, gkey = ctx.unpack, ('dgcat', id(mod_w))
            flat = st is not None and all(getattr(q, '_rcfd_flat', False) and q.grad is not None
                                          for q in (bw.weight, bw.bias, bp.weight, bp.bias))
            if flat:
                gbuf = st['bufs'].get(gkey)
                if gbuf is None:
                    gbuf = st['bufs'][gkey] = torch.empty(2, 2 * c, device=ctx.device, dtype=torch.float32)
                    st['dirty'] = True
                dg, db = gbuf[0], gbuf[1]
                gdeferred = st['batch'] is not None and gkey in st['batch']['keys']
            else:
                dg = torch.empty(2 * c, device=ctx.device, dtype=torch.float32)
                db = torch.empty_like(dg)
                gdeferred = False
            dy = ops.bn_act_bwd(dzy, y, scale, shift, mean, invstd, ACT_NONE, dg, db, sums=ctx.zeroed_sums(2 * c))
            for bn, lo in ((bw, 0), (bp, c)):
                if gdeferred:
                    tape.param_grads.append((bn.weight, bn.weight.grad))
                    tape.param_grads.append((bn.bias, bn.bias.grad))
                    continue
                tape.param_grads.append((bn.weight, dg[lo:lo + c]))
                tape.param_grads.append((bn.bias, db[lo:lo + c]))
                if flat and len(st['items'].get(gkey, ())) < 4:
                    st['items'].setdefault(gkey, []).extend([
                        (ops.COPY_F32, gbuf, lo, bn.weight.grad, bn.weight, dict(total=c)),
                        (ops.COPY_F32, gbuf, 2 * c + lo, bn.bias.grad, bn.bias, dict(total=c))])
            ukey = ('dwcat', id(mod_w))
            gs = [_grad_dst(ww), _grad_dst(wp)]
            buf, deferred = _wgrad_slot(ctx, ukey, (2 * c, 1, dep.shape[3]), ww, gs[0])
            if buf is not None and (gs[1] is not wp.grad or not getattr(wp, '_rcfd_flat', False)):
                buf, deferred = None, False
            dw = ops.conv2d_wgrad(dep, dy, 1, 1, engine=ctx.engine, x3=ctx.x3, out=buf)              # [2c, 1, cd]
            for (wparam, lo), g in zip(((ww, 0), (wp, c)), gs):
                if not deferred:
                    ops.unpack_wgrad(dw[lo:lo + c], g)
                    if buf is not None and len(ctx.unpack['items'].get(ukey, ())) < 2:
                        _unpack_conv(ctx, ukey, dw, g, wparam, src_off=lo * dw.shape[2])
                tape.param_grads.append((wparam, g))
            wd = ctx.packed(('wcatd', id(mod_w)), [ww, wp],
                            lambda: ops.pack_weight(torch.cat([ww.detach(), wp.detach()], 0), dtype, dgrad=True),
                            spec=lambda: ops.spec_pack_stacked_1x1([ww.detach(), wp.detach()], dtype, dgrad=True))
            tape.add_grad(dep, ops.conv2d(dy, wd, dep.shape[3], 1, 1, engine=ctx.engine))
        tape.add_step(bwd)
    return out


def max_pool(ctx, x):
    vec = 8 if x.dtype == torch.bfloat16 else 4
    if ctx.tape is not None and x.shape[3] % vec == 0 and hasattr(ops, 'maxpool3x3s2_idx'):
        # training: record the arg-max positions (1 byte / element) so that the backward does not re-scan the input
        out, idx = ops.maxpool3x3s2_idx(x)
        tape = ctx.tape
        hw = (x.shape[1], x.shape[2])

        def bwd():
            d = tape.grad_of(out)
            if d is not None:
                tape.add_grad(x, ops.maxpool3x3s2_bwd_idx(d, idx, hw))
        tape.add_step(bwd)
        return out
    out = ops.maxpool3x3s2(x)
    if ctx.tape is not None:
        tape = ctx.tape

        def bwd():
            d = tape.grad_of(out)
            if d is not None:
                tape.add_grad(x, ops.maxpool3x3s2_bwd(x, d))
        tape.add_step(bwd)
    return out


def res_block(ctx, blk, x):
    """net_utils.ResNetBlock (reference src/net_utils.py:309-323)."""
    c1 = conv_unit(ctx, blk.conv1, x)
    if blk.stride != 1 or blk.in_channels != blk.out_channels:
        sc = conv_unit(ctx, blk.projection, x)
    else:
        sc = x
    return conv_unit(ctx, blk.conv2, c1, residual=sc)


def res_stage(ctx, stage, x):
    for blk in stage:
        x = res_block(ctx, blk, x)
    return x


def decoder_block(ctx, blk, x, skip, shape):
    """net_utils.DecoderBlock (reference src/net_utils.py:535-569)."""
    if skip is not None:
        shape = (skip.shape[1], skip.shape[2])
    elif shape is None:
        shape = (2 * x.shape[1], 2 * x.shape[2])
    d = conv_unit(ctx, blk.deconv.conv, x, in_size=(int(shape[0]), int(shape[1])))
    return conv_unit(ctx, blk.conv, d, x1=skip if blk.skip_channels > 0 else None)


# ----------------------------------------------------------------------------- whole graphs
def stem_input(ctx, x_nchw):
    """NCHW float API tensor -> what the 7x7/s2 stem consumes: a space-to-depth NHWC tensor on the bf16 fast
    path (even H, W), else NHWC with channels zero-padded to CPAD.  Returns (tensor, is_s2d)."""
    h, w = x_nchw.shape[-2:]
    if (ctx.dtype == torch.bfloat16 or ctx.x3) and h % 2 == 0 and w % 2 == 0 and 4 * x_nchw.shape[1] <= CPAD:
        return ops.nchw_to_s2d(x_nchw.float(), ctx.dtype, CPAD), True
    return ops.nchw_to_nhwc(x_nchw.float(), ctx.dtype, cpad=CPAD), False


def fusionnet_encoder(ctx, enc, image, depth, stem_s2d=False):
    """networks.FusionNetEncoder.forward (reference src/networks.py:840-1005); NHWC in / out.
    ``image`` / ``depth`` may be callables producing the stem inputs (so the layout conversion of each
    branch is issued on that branch's stream).  Multi-stream schedule: the image branch runs on MAIN,
    the depth branch on DEPTH, every gated fusion level on FUSE after both of its inputs."""
    ctx.fork()
    ctx.prepack()
    with ctx.on(MAIN):
        if callable(image):
            image, stem_s2d = image()
        ci = conv_unit(ctx, enc.conv1_image, image, want_input_grad=False, stem_s2d=stem_s2d)
    with ctx.on(DEPTH):
        if callable(depth):
            depth, s2d_d = depth()
            assert s2d_d == stem_s2d
        cd = conv_unit(ctx, enc.conv1_depth, depth, want_input_grad=False, stem_s2d=stem_s2d)
    ctx.wait(FUSE, (MAIN, DEPTH))
    with ctx.on(FUSE):
        layers = [fusion_level(ctx, enc.conv1_weight, enc.conv1_project, cd, ci)]
    with ctx.on(MAIN):
        xi = max_pool(ctx, ci)
    with ctx.on(DEPTH):
        xd = max_pool(ctx, cd)
    for level in range(2, 8):
        si = getattr(enc, 'blocks%d_image' % level)
        if si is None:
            break
        if level == 5 and ctx.tape is not None:
            ctx.tape.mark_split()          # levels 5, 6 and the decoder = the last 75 % of the parameters (registration order)
        with ctx.on(MAIN):
            xi = res_stage(ctx, si, xi)
        with ctx.on(DEPTH):
            xd = res_stage(ctx, getattr(enc, 'blocks%d_depth' % level), xd)
        ctx.wait(FUSE, (MAIN, DEPTH))
        with ctx.on(FUSE):
            layers.append(fusion_level(ctx, getattr(enc, 'conv%d_weight' % level),
                                       getattr(enc, 'conv%d_project' % level), xd, xi))
    ctx.join()
    return layers[-1], layers[:-1]


def resnet_encoder(ctx, enc, x, stem_s2d=False):
    """networks.ResNetEncoder.forward (reference src/networks.py:232-268)."""
    layers = [conv_unit(ctx, enc.conv1, x, want_input_grad=False, stem_s2d=stem_s2d)]
    y = max_pool(ctx, layers[-1])
    for level in range(2, 8):
        stage = getattr(enc, 'blocks%d' % level)
        if stage is None:
            break
        y = res_stage(ctx, stage, y)
        layers.append(y)
    return layers[-1], layers[:-1]


def _logit_output(ctx, mod, x, head):
    """An intermediate output of the multi-resolution decoder: 3x3 conv -> float logits [N, h, w, 1], their depth
    (bounded head, when ``head`` is given) and the 2x bilinear up-sampling that joins the next skip.  Returns
    (logits, depth or None, up); with a tape the gradients of depth and up accumulate on the logits."""
    logits = ops.conv2d(x, ctx.weight(mod), mod.out_channels, 3, 1, out_f32=True, engine=ctx.engine)
    if ctx.tape is not None:
        # d(logits): float, 1 channel -> the compute dtype with CPAD_DY channels (rows of >= 32 bytes for the conv engines)
        to_dy = lambda d: ops.nchw_to_nhwc(d.view(d.shape[0], 1, d.shape[1], d.shape[2]), ctx.dtype,
                                           cpad=CPAD_DY if (ctx.dtype == torch.bfloat16 or ctx.x3) else 8)
        _record_conv_backward(ctx, mod, x, None, None, logits, None, True, pre=to_dy)
    depth = None
    if head is not None:
        depth = ops.epilogue_f32(logits, None, None, ACT_DEPTH_HEAD, act_params=head)
    up = ops.bilinear2x(logits)
    if ctx.tape is not None:
        tape = ctx.tape

        def bwd():
            for t, back in ((up, ops.bilinear2x_bwd),
                            (depth, (lambda d: ops.depth_head_bwd(d, depth, head[0], head[1], torch.float32, cpad=1)))):
                if t is None:
                    continue
                d = tape.grad_of(t)
                if d is not None:
                    tape.add_grad(logits, back(d.contiguous()))
        tape.add_step(bwd)
    return logits, depth, up


def _skip_with_logit(ctx, skip, up):
    """torch.cat([skip, up], 1) of the reference (src/networks.py:1608, 1624, 1640), channels padded to a multiple of 16."""
    cat = ops.concat_logit(skip, up, ctx.dtype)
    if ctx.tape is not None:
        tape = ctx.tape
        c = 0 if skip is None else skip.shape[3]

        def bwd():
            d = tape.grad_of(cat)
            if d is None:
                return
            dskip, dup = ops.split_logit(d, c)
            if skip is not None:
                tape.add_grad(skip, dskip)
            tape.add_grad(up, dup)
        tape.add_step(bwd)
    return cat


def multiscale_decoder(ctx, dec, latent, skips, shape, head=None):
    """networks.MultiScaleDecoder.forward (reference src/networks.py:1557-1657).  Returns (output, last feature map);
    with n_resolution > 1 the coarser outputs (depth when ``head`` is given, else logits; float [N, h, w, 1], coarsest
    first) are left in ``ctx.multiscale``."""
    x = latent
    n = len(skips) - 1
    up = None
    ctx.multiscale = []
    for b in range(dec.n_blocks - 1, -1, -1):
        blk = getattr(dec, 'deconv%d' % b)
        skip = skips[n] if n >= 0 else None
        n -= 1
        if up is not None:
            skip = _skip_with_logit(ctx, skip, up)
        x = decoder_block(ctx, blk, x, skip, shape if skip is None else None)
        up = None
        if ctx.taps is not None:
            ctx.taps['deconv%d' % b] = x
        out_mod = getattr(dec, 'output%d' % b, None) if b >= 1 else None
        if out_mod is not None:
            logits, depth, up = _logit_output(ctx, out_mod, x, head)
            ctx.multiscale.append(depth if head is not None else logits)
    out = conv_unit(ctx, dec.output0, x, head=head) if head is not None else \
        ops.conv2d(x, ctx.weight(dec.output0), dec.output0.out_channels, 3, 1, out_f32=True, engine=ctx.engine)
    return out, x


def finish_bn_counters(ctx):
    if ctx.bn_counters:
        torch._foreach_add_(ctx.bn_counters, 1)
        ctx.bn_counters = []


# ----------------------------------------------------------------------------- standalone module calls
def _ctx_for(mod, x):
    training = False      # standalone block calls are inference-only (training goes through the model wrappers)
    return Context(torch.float32, training, x.device)


def standalone_activation(x, kind):
    flat = x.contiguous().view(-1, 4) if x.numel() % 4 == 0 else None
    if flat is None:
        raise ValueError('standalone activation needs numel % 4 == 0')
    return ops.bn_act(flat, None, None, _ACT[kind]).view(x.shape)


def standalone(mod, kind, x, **kw):
    """Run one module of the tree by itself on NCHW float tensors (inference semantics of
    ``mod.training`` is NOT honoured for BatchNorm statistics: eval-mode folding is used)."""
    ctx = _ctx_for(mod, x)
    to = lambda t: ops.nchw_to_nhwc(t.float(), ctx.dtype, cpad=CPAD if t.shape[1] < CPAD else None)
    back = ops.nhwc_to_nchw
    if kind == 'conv':
        return back(conv_unit(ctx, mod, to(x)))
    if kind == 'upconv':
        return back(conv_unit(ctx, mod.conv, to(x), in_size=tuple(kw['shape'])))
    if kind == 'resblock':
        return back(res_block(ctx, mod, to(x)))
    if kind == 'decoder_block':
        skip = kw.get('skip')
        return back(decoder_block(ctx, mod, to(x), to(skip) if skip is not None else None, kw.get('shape')))
    if kind == 'resnet_encoder':
        latent, skips = resnet_encoder(ctx, mod, to(x))
        return back(latent), [back(s) for s in skips]
    if kind == 'fusionnet_encoder':
        latent, skips = fusionnet_encoder(ctx, mod, to(x), to(kw['depth']))
        return back(latent), [back(s) for s in skips]
    if kind == 'decoder':
        out, _ = multiscale_decoder(ctx, mod, to(x), [to(s) for s in kw['skips']], kw.get('shape'))
        return [out.view(out.shape[0], 1, out.shape[1], out.shape[2])]
    if kind == 'radarnet_encoder':
        latent, skips = radarnet_encoder(ctx, mod, to(x), kw['points'], kw['boxes'])
        return back(latent), [back(s) for s in skips]
    raise ValueError(kind)


# ----------------------------------------------------------------------------- RadarNet column
def boxes_to_rois(boxes_list, device):
    rows = []
    for b, boxes in enumerate(boxes_list):
        boxes = boxes.to(device=device, dtype=torch.float32)
        rows.append(torch.cat([torch.full((boxes.shape[0], 1), float(b), device=device), boxes], dim=1))
    return torch.cat(rows, dim=0).contiguous()


def _roi_pool(ctx, feat, rois, out_size, scale):
    out = ops.roi_pool(feat, rois, out_size, scale)
    if ctx.tape is not None:
        tape = ctx.tape

        def bwd():
            d = tape.grad_of(out)
            if d is not None:
                tape.add_grad(feat, ops.roi_pool_bwd(feat, rois, d, out_size, scale))
        tape.add_step(bwd)
    return out


def radarnet_encoder(ctx, enc, image, points, boxes_list, stem_s2d=False):
    """networks.RadarNetV1Encoder.forward (reference src/networks.py:1203-1256); image NHWC.  With a tape (training,
    reference src/radarnet_main.py:392-399) every piece records its backward: roi_pool -> arg-max scatter, the point MLP ->
    rcfd_linear_leaky_bwd per layer, the latent concat -> channel split."""
    ph, pw = enc.input_patch_size_image
    lat_h, lat_w = int(ph // 32.0), int(pw // 32.0)
    scales = [1 / 2.0, 1 / 4.0, 1 / 8.0, 1 / 16.0, 1 / 32.0, 1 / 64.0, 1 / 128.0]
    latent_img, skips_img = resnet_encoder(ctx, enc.encoder_image, image, stem_s2d=stem_s2d)
    rois = boxes_to_rois(boxes_list, image.device)
    latent_pooled = _roi_pool(ctx, latent_img, rois, (lat_h, lat_w), 1 / 32.0)
    skips = [_roi_pool(ctx, s, rois, (int(ph * scales[i]), int(pw * scales[i])), scales[i])
             for i, s in enumerate(skips_img)]
    x = points.to(device=image.device, dtype=torch.float32).contiguous()
    acts = [x]
    for fc in enc.encoder_depth.mlp:
        x = ops.linear_leaky(x, fc.fully_connected.weight.detach(), fc.fully_connected.bias.detach())
        acts.append(x)
    k = x.shape[0]
    cl = enc.n_neuron_latent_depth
    # reference views the MLP output as (K, C, h, w) NCHW (src/networks.py:1252) -> NHWC
    lat_d = ops.nchw_to_nhwc(x.view(k, cl, lat_h, lat_w), ctx.dtype)
    latent = torch.cat([latent_pooled, lat_d], dim=3).contiguous()      # channel concat of two small tensors
    if ctx.tape is not None:
        tape = ctx.tape
        c_img = latent_pooled.shape[3]
        layers = list(enc.encoder_depth.mlp)

        def bwd():
            d = tape.grad_of(latent)
            if d is None:
                return
            tape.add_grad(latent_pooled, d[..., :c_img].contiguous())
            g = ops.nhwc_to_nchw(d[..., c_img:].contiguous()).view(k, -1)          # float32, the MLP's (K, C*h*w) order
            for i in range(len(layers) - 1, -1, -1):
                fc = layers[i].fully_connected
                g, dw, db = ops.linear_leaky_bwd(acts[i], fc.weight.detach(), acts[i + 1], g, want_dx=i > 0)
                tape.param_grads.append((fc.weight, dw))
                tape.param_grads.append((fc.bias, db))
        tape.add_step(bwd)
    return latent, skips
