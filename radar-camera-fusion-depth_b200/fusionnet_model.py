"""
FusionNetModel with the reference's constructor / method surface (reference:
src/fusionnet_model.py:7-401), executed on librcfd_b200.so.

Tensors at the API are NCHW float32 like the reference; parameters keep the reference's
names and OIHW float32 layout so ``state_dict`` / checkpoints / torch.optim interoperate.
``forward`` returns a tensor that supports ``loss.backward()``: the whole network is one
autograd node whose backward replays the engine's tape (dgrad / wgrad / BN backward
kernels) and fills ``parameter.grad``.

Extras that the reference does not have (all optional):
  * ``set_precision('fp32' | 'bf16' | 'bf16x3' | 'bf16x6')`` -- arithmetic inside the kernels
    (accumulation is always fp32).  'fp32': fp32 storage, SIMT FMA convolutions.  'bf16': bf16 storage,
    tcgen05 engines (the fast path).  'bf16x3' / 'bf16x6': the tensor-core PARITY modes -- fp32 storage,
    every convolution / weight gradient runs as 3 / 6 passes of the same tcgen05 engines over 2- / 3-part
    bf16 splits of the fp32 operands (~2^-16 / ~2^-23 per product: 'bf16x6' is fp32-class; include/rcfd.h).
  * ``forward(..., return_logits=True)`` for tests.
"""
import os

import torch

import fusionnet_losses as losses
import networks
from rcfd import engine, ops


class _FusionNetFunction(torch.autograd.Function):
    """One autograd node for the whole encoder-decoder."""

    @staticmethod
    def forward(fctx, model, record, image, input_depth, *params):
        out_nhwc, ectx = model._run(image, input_depth, record=record)
        # multi-resolution decoder: the coarser outputs first, the full-resolution one last (reference :160-170)
        outs = list(getattr(ectx, 'multiscale', [])) + [out_nhwc]
        fctx.model, fctx.ectx, fctx.outs_nhwc = model, ectx, outs
        fctx.set_materialize_grads(False)
        views = tuple(t.view(t.shape[0], 1, t.shape[1], t.shape[2]) for t in outs)        # C == 1: NHWC and NCHW coincide
        return views if len(views) > 1 else views[0]

    @staticmethod
    def backward(fctx, *grad_outs):
        ectx = fctx.ectx
        n_in = 4 + len(list(fctx.model.parameters()))
        if all(g is None for g in grad_outs) or ectx is None or ectx.tape is None:
            return (None,) * n_in
        tape = ectx.tape
        for t, g in zip(fctx.outs_nhwc, grad_outs):
            if g is not None:
                tape.set_grad(t, g.contiguous().float().view(t.shape))
        tape.backward()
        fctx.model._deliver_grads(tape.param_grads)
        tape.param_grads = []
        fctx.ectx = None
        return (None,) * n_in


class FusionNetModel(object):
    """Image + radar depth fusion network (reference src/fusionnet_model.py:7)."""

    def __init__(self, input_channels_image, input_channels_depth, encoder_type, n_filters_encoder_image,
                 n_filters_encoder_depth, fusion_type, decoder_type, n_resolution_decoder, n_filters_decoder,
                 deconv_type, activation_func, weight_initializer, min_predict_depth, max_predict_depth,
                 device=torch.device('cuda')):
        self.encoder_type = encoder_type
        self.min_predict_depth = min_predict_depth
        self.max_predict_depth = max_predict_depth
        self.device = device
        self.compute_dtype = torch.float32
        self.x3 = 0
        self.precision = 'fp32'
        self.conv_engine = ops.ENGINE_AUTO
        # image branch / depth branch / fusion / weight gradients on parallel CUDA streams (RCFD_MULTISTREAM=0: one stream)
        self.multistream = os.environ.get('RCFD_MULTISTREAM', '1') != '0'
        self._cache = {}
        self.grad_hook = None          # called with the list of (param, grad) after backward (DDP)

        if fusion_type not in ('add', 'weight', 'weight_and_project', 'concat'):
            raise ValueError('Unsupported fusion type: {}'.format(fusion_type))
        if 'fusionnet18' in encoder_type or 'resnet18' in encoder_type:
            n_layer = 18
        elif 'fusionnet34' in encoder_type or 'resnet34' in encoder_type:
            n_layer = 34
        else:
            raise ValueError('Unsupported encoder type: {}'.format(encoder_type))
        if not ('fusionnet18' in encoder_type or 'fusionnet34' in encoder_type):
            raise ValueError('Unsupported encoder type on the B200 path: {} (image-only resnet encoders are not '
                             'used by any shipped FusionNet config)'.format(encoder_type))
        self.encoder = networks.FusionNetEncoder(
            n_layer=n_layer, input_channels_image=input_channels_image, input_channels_depth=input_channels_depth,
            n_filters_encoder_image=n_filters_encoder_image, n_filters_encoder_depth=n_filters_encoder_depth,
            weight_initializer=weight_initializer, activation_func=activation_func,
            use_batch_norm='batch_norm' in encoder_type, fusion_type=fusion_type)
        n_filters_encoder = list(n_filters_encoder_image)
        n_skips = n_filters_encoder[:-1][::-1] + [0]
        if 'multiscale' in decoder_type:
            self.decoder = networks.MultiScaleDecoder(
                input_channels=n_filters_encoder[-1], output_channels=1, n_resolution=n_resolution_decoder,
                n_filters=n_filters_decoder, n_skips=n_skips, weight_initializer=weight_initializer,
                activation_func=activation_func, output_func='linear', use_batch_norm='batch_norm' in decoder_type,
                deconv_type=deconv_type)
        else:
            raise ValueError('Unsuported decoder type: {}'.format(decoder_type))
        if not ('batch_norm' in encoder_type and 'batch_norm' in decoder_type):
            raise ValueError('the B200 path implements the shipped batch_norm encoder / decoder variants')
        self.to(self.device)

    # ------------------------------------------------------------------ precision / engine knobs
    def set_precision(self, precision):
        self.compute_dtype = {'fp32': torch.float32, 'bf16': torch.bfloat16, 'bf16x3': torch.float32, 'bf16x6': torch.float32}[precision]
        self.x3 = {'bf16x3': 2, 'bf16x6': 3}.get(precision, 0)
        self.precision = precision
        self._invalidate()
        return self

    def _invalidate(self):
        """Drop everything derived from the parameters: packed weights, folded BatchNorm, and the captured CUDA
        graphs (a graph holds raw pointers to the packed tensors of the pass it was captured from)."""
        self._cache.clear()
        self._graphs = {}
        self._train_graphs = {}

    # ------------------------------------------------------------------ execution
    def _run(self, image, input_depth, record=False, return_logits=False, taps=None):
        # rcfd.ops refuses non-CUDA tensors: there is no CPU fallback behind this call
        with ops.hold_allocations():
            ectx = engine.Context(self.compute_dtype, self.encoder.training, image.device, cache=self._cache,
                                  record=record, engine=self.conv_engine, multistream=self.multistream, x3=self.x3,
                                  external_pack=getattr(self, '_external_pack', False))
            ectx.taps = taps
            ectx.grad_split = getattr(self, '_grad_split_addr', None) if record else None
            # the layout conversion of each input is issued on its branch's stream
            latent, skips = engine.fusionnet_encoder(ectx, self.encoder, lambda: engine.stem_input(ectx, image),
                                                     lambda: engine.stem_input(ectx, input_depth))
            if taps is not None:
                taps['latent'] = latent
                for i, s in enumerate(skips):
                    taps['skip%d' % (i + 1)] = s
            head = None if return_logits else (float(self.min_predict_depth),
                                               float(self.min_predict_depth) / float(self.max_predict_depth))
            out, _ = engine.multiscale_decoder(ectx, self.decoder, latent, skips, image.shape[-2:], head=head)
            engine.finish_bn_counters(ectx)
        return out, ectx

    def forward(self, image, input_depth, return_multiscale=False, return_logits=False):
        """N x 3 x H x W image, N x 2 x H x W (depth, response) -> N x 1 x H x W depth in
        [min, max] metres (reference src/fusionnet_model.py:140-170)."""
        if return_logits:
            out, ectx = self._run(image, input_depth, return_logits=True)
            outs = [t.view(t.shape[0], 1, t.shape[1], t.shape[2]) for t in list(ectx.multiscale) + [out]]
        else:
            params = self.parameters()
            # grad mode is off inside autograd.Function.forward, so decide here whether to tape
            record = torch.is_grad_enabled() and self.encoder.training and any(p.requires_grad for p in params)
            outs = _FusionNetFunction.apply(self, record, image, input_depth, *params)
            outs = list(outs) if isinstance(outs, tuple) else [outs]
        return outs if return_multiscale else outs[-1]

    def _feed(self, entry, static, tensors):
        """Bring this step's inputs into the graph's static buffers.  Device tensors: one device copy each.
        Host tensors (pinned): the host->device copies run on a COPY stream into one of two staging sets, so the
        copy of step i+1 overlaps the compute of step i (the host enqueues step i+1 while step i is running);
        the graph's stream then only does a device-to-device copy (staging -> static) before the replay."""
        if all(t.is_cuda for t in tensors):
            for s, t in zip(static, tensors):
                s.copy_(t, non_blocking=True)
            return
        main = torch.cuda.current_stream()
        if 'stage' not in entry:
            entry['stage'] = [[torch.empty_like(s) for s in static] for _ in range(2)]
            entry['free'] = [torch.cuda.Event(), torch.cuda.Event()]
            entry['copy_stream'] = torch.cuda.Stream()
            entry['n'] = 0
            for ev in entry['free']:
                ev.record(main)
        slot = entry['n'] % 2
        entry['n'] += 1
        cs = entry['copy_stream']
        cs.wait_event(entry['free'][slot])              # the step that consumed this staging set has read it
        with torch.cuda.stream(cs):
            for s, t in zip(entry['stage'][slot], tensors):
                s.copy_(t, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(cs)
        main.wait_event(ready)
        for s, g in zip(static, entry['stage'][slot]):
            s.copy_(g, non_blocking=True)
        entry['free'][slot].record(main)

    def forward_graphed(self, image, input_depth):
        """Inference forward replayed from a CUDA graph (captured once per input shape / precision):
        the ~150 kernel launches of one forward become one graph launch, so small batches are not
        launch-bound.  eval() + no_grad semantics; returns a tensor owned by the graph (overwritten
        by the next call)."""
        if self.encoder.training:
            raise RuntimeError('forward_graphed is inference-only: call model.eval() first')
        # eval graphs hold the packed weights / folded BatchNorm of the pass they were captured from: the key carries
        # the parameter epoch (rcfd.optim.FusedAdam steps) and torch's version counters (optimizers, load_state_dict)
        key = (tuple(image.shape), tuple(input_depth.shape), self.precision, self.conv_engine, self.multistream,
               engine._PARAM_EPOCH[0], sum(t._version for t in self._state_tensors()))
        entry = self._graphs.get(key) if hasattr(self, '_graphs') else None
        if entry is None:
            if not hasattr(self, '_graphs'):
                self._graphs = {}
            dev = next(self.encoder.parameters()).device
            s_img = torch.empty(tuple(image.shape), device=dev, dtype=torch.float32)
            s_dep = torch.empty(tuple(input_depth.shape), device=dev, dtype=torch.float32)
            s_img.copy_(image)
            s_dep.copy_(input_depth)
            with torch.no_grad():
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    for _ in range(2):                      # warm-up: pack weights, fold BN, set kernel attributes
                        self.forward(s_img, s_dep)
                torch.cuda.current_stream().wait_stream(side)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=engine.capture_stream(dev)):
                    s_out = self.forward(s_img, s_dep)
            entry = {'graph': graph, 'static': [s_img, s_dep], 'out': s_out}
            for k in [k for k in self._graphs if k[:5] == key[:5]]:      # same shape / mode, older parameters
                del self._graphs[k]
            self._graphs[key] = entry
        self._feed(entry, entry['static'], (image, input_depth))
        entry['graph'].replay()
        return entry['out']

    def forward_graphed_raw(self, raw, response_multiplier=256.0):
        """forward_graphed fed with a batch in the reference's on-disk sample types: ``raw`` = (image uint8 N x H x W x 3,
        depth uint16 N x H x W, response uint16 N x H x W; int16-viewed tensors are fine), pinned host or device tensors:
        7 bytes per pixel cross PCIe instead of 20; value codec (/ 255, / 256, <= 0 -> 0) and layout change on the device
        (rcfd_decode_crop) straight into the graph's input buffers."""
        n, h, w, _ = raw[0].shape
        dev = next(self.encoder.parameters()).device
        if not hasattr(self, '_raw_eval'):
            self._raw_eval = {}
        key = (n, h, w)
        st = self._raw_eval.get(key)
        if st is None:
            st = {'image': torch.zeros(n, 3, h, w, device=dev), 'depth': torch.zeros(n, 2, h, w, device=dev),
                  'stage': [[torch.empty(t.shape, dtype=t.dtype, device=dev) for t in raw[:3]] for _ in range(2)],
                  'free': [torch.cuda.Event(), torch.cuda.Event()], 'copy': torch.cuda.Stream(), 'n': 0}
            for ev in st['free']:
                ev.record(torch.cuda.current_stream())
            self._raw_eval[key] = st
        src = raw
        if not all(t.is_cuda for t in raw[:3]):
            main = torch.cuda.current_stream()
            slot = st['n'] % 2
            st['n'] += 1
            st['copy'].wait_event(st['free'][slot])
            with torch.cuda.stream(st['copy']):
                for s_, t in zip(st['stage'][slot], raw[:3]):
                    s_.copy_(t, non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(st['copy'])
            main.wait_event(ready)
            src = st['stage'][slot]
        ops.decode_crop(src[0], 255.0, out=st['image'])
        ops.decode_crop(src[1], 256.0, out=st['depth'], out_channel=0)
        ops.decode_crop(src[2], response_multiplier, out=st['depth'], out_channel=1)
        if src is not raw:
            st['free'][slot].record(torch.cuda.current_stream())
        return self.forward_graphed(st['image'], st['depth'])

    def train_step_graphed(self, image, input_depth, ground_truth, lidar_map, optimizer, w_lidar_loss,
                           outlier_removal=None):
        """One optimisation step of the canonical configuration (reference src/fusionnet_main.py:366-399:
        forward -> ground-truth outlier removal -> masked L1 (+ lidar term) -> backward -> Adam) with the
        ~480 kernel launches of forward + loss + backward replayed from ONE CUDA graph (captured once per
        input shape / precision); the gradient all-reduce (data parallel), the optimiser step and the batched
        weight packing of the next step follow eagerly, so learning-rate schedules and the Adam step count stay
        host-side.  Same arithmetic as ``forward`` / ``compute_loss`` / ``loss.backward()`` / ``optimizer.step()``;
        returns the loss as a 0-d tensor owned by the graph (overwritten by the next call)."""
        tensors = (image, input_depth, ground_truth, lidar_map)
        entry = self._train_graph_entry([tuple(t.shape) for t in tensors], optimizer, w_lidar_loss, outlier_removal,
                                        fill=lambda static: [s.copy_(t) for s, t in zip(static, tensors)])
        self._feed(entry, entry['static'], tensors)
        return self._replay_train(entry, optimizer)

    def train_step_graphed_raw(self, raw, optimizer, w_lidar_loss, outlier_removal=None, response_multiplier=256.0):
        """The same step fed with a batch in the reference's ON-DISK sample types (what rcfd.data.FusionNetRawDataset
        yields): ``raw`` = (image uint8 N x H x W x 3, depth / response / ground truth / lidar uint16 N x H x W,
        int16-viewed tensors are fine), pinned host or device tensors.  11 bytes per pixel cross PCIe instead of the 28
        of five float32 tensors; the value codec of src/data_utils.py:167-198, 238-318 (/ 255 for the image with
        normalized_image_range [0, 1], / 256 for the maps, <= 0 -> 0) and the HWC -> CHW change run on the device
        (rcfd_decode_crop) straight into the graph's input buffers; host copies go through a double-buffered staging
        set on a copy stream like ``_feed``."""
        n, h, w, _ = raw[0].shape
        shapes = [(n, 3, h, w), (n, 2, h, w), (n, 1, h, w), (n, 1, h, w)]
        dev = next(self.encoder.parameters()).device

        def decode(static, src):
            ops.decode_crop(src[0], 255.0, out=static[0])
            ops.decode_crop(src[1], 256.0, out=static[1], out_channel=0)
            ops.decode_crop(src[2], response_multiplier, out=static[1], out_channel=1)
            ops.decode_crop(src[3], 256.0, out=static[2])
            ops.decode_crop(src[4], 256.0, out=static[3])
        entry = self._train_graph_entry(shapes, optimizer, w_lidar_loss, outlier_removal,
                                        fill=lambda static: decode(static, [t.to(dev) for t in raw[:5]]))
        if all(t.is_cuda for t in raw[:5]):
            decode(entry['static'], raw)
        else:
            main = torch.cuda.current_stream()
            if 'raw_stage' not in entry:
                entry['raw_stage'] = [[torch.empty(t.shape, dtype=t.dtype, device=dev) for t in raw[:5]] for _ in range(2)]
                entry['raw_free'] = [torch.cuda.Event(), torch.cuda.Event()]
                entry['raw_copy_stream'] = torch.cuda.Stream()
                entry['raw_n'] = 0
                for ev in entry['raw_free']:
                    ev.record(main)
            slot = entry['raw_n'] % 2
            entry['raw_n'] += 1
            cs = entry['raw_copy_stream']
            cs.wait_event(entry['raw_free'][slot])          # the step that decoded this staging set has read it
            with torch.cuda.stream(cs):
                for s_, t in zip(entry['raw_stage'][slot], raw[:5]):
                    s_.copy_(t, non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(cs)
            main.wait_event(ready)
            decode(entry['static'], entry['raw_stage'][slot])
            entry['raw_free'][slot].record(main)
        return self._replay_train(entry, optimizer)

    def _replay_train(self, entry, optimizer):
        entry['graph'].replay()
        if entry.get('graph_b') is not None:
            # data parallel: the gradients of encoder levels >= 5 and of the decoder (the tail of the flat buffer, 75 % of
            # the bytes) are final after the first graph; their all-reduce runs on NCCL's stream while the second graph
            # computes the rest of the backward
            self.grad_hook.reduce_slice(entry['split_off'], None)
            entry['graph_b'].replay()
            self.grad_hook.reduce_slice(0, entry['split_off'])
            self.grad_hook.finish()
        elif self.grad_hook is not None:          # gradients live in the optimiser's flat buffer (written by the graph)
            self.grad_hook(entry['grads'])
        optimizer.step()
        if entry.get('pack') is not None:
            entry['pack']['table'].run()            # next step's packed weights (the graph reads the persistent buffers)
        return entry['loss']

    def _train_graph_entry(self, shapes, optimizer, w_lidar_loss, outlier_removal, fill):
        """The captured step graph for these input shapes (built on first use; ``fill(static)`` puts a first batch into
        the graph's input buffers for the warm-up pass)."""
        if not self.encoder.training:
            raise RuntimeError('train_step_graphed needs model.train()')
        if not w_lidar_loss > 0.0:
            raise ValueError('train_step_graphed implements the canonical loss (l1, w_lidar_loss > 0)')
        if not all(getattr(p, '_rcfd_flat', False) for p in self.parameters()):
            raise RuntimeError('train_step_graphed needs rcfd.optim.FusedAdam (gradients written in place into its flat buffer)')
        if not hasattr(self, '_train_graphs'):
            self._train_graphs = {}
        # overlapped gradient all-reduce (rcfd.parallel.use_flat_gradients): split the step into two graphs.  Opt-in
        # (RCFD_DDP_OVERLAP=1): measured on 2 and 8 B200s it changes nothing (5.867 vs 5.853 ms, 5.967 vs 5.970 ms per step):
        # the ~0.26 ms an 8-GPU step costs over a 1-GPU step is not hidden by running the all-reduce beside the backward
        split_off = None
        hook = self.grad_hook
        if (hook is not None and getattr(hook, 'flat_grad', None) is not None and hasattr(hook, 'reduce_slice')
                and self.multistream and not self.x3 and os.environ.get('RCFD_DDP_OVERLAP', '0') == '1'):
            first = getattr(self.encoder, 'blocks5_image', None)
            first = next(first.parameters(), None) if first is not None else None
            if first is not None and getattr(first, '_rcfd_flat', False) and first.grad is not None:
                split_off = (first.grad.data_ptr() - hook.flat_grad.data_ptr()) // 4
        key = (tuple(shapes[0]), tuple(shapes[1]), self.precision, self.conv_engine, float(w_lidar_loss),
               None if outlier_removal is None else (outlier_removal.kernel_size, outlier_removal.threshold),
               id(optimizer), self.multistream, split_off)
        entry = self._train_graphs.get(key)
        if entry is None:
            dev = next(self.encoder.parameters()).device
            static = [torch.empty(tuple(sh), device=dev, dtype=torch.float32) for sh in shapes]
            fill(static)

            self._grad_split_addr = None if split_off is None else hook.flat_grad.data_ptr() + 4 * split_off
            state = {}

            def body(part=None):
                """part None: the whole step; 0: forward + loss + backward down to the split marker; 1: the rest."""
                if part in (None, 0):
                    out, ectx = self._run(static[0], static[1], record=True)
                    n, h, w, _ = out.shape
                    gt = static[2] if outlier_removal is None else outlier_removal.remove_outliers(static[2])
                    loss, dout = ops.masked_l1_loss(out.view(n, 1, h, w), gt, static[3], float(w_lidar_loss), want_grad=True)
                    tape = ectx.tape
                    tape.set_grad(out, dout.view(out.shape))
                    state['tape'], state['loss'] = tape, loss
                    if part == 0:
                        tape.backward(part=0)
                        return
                    tape.backward()
                else:
                    tape, loss = state['tape'], state['loss']
                    tape.backward(part=1)
                grads, tape.param_grads = tape.param_grads, []
                self._deliver_grads(grads, hook=False)       # the few gradients not written in place (captured copies)
                return loss.view(()), grads

            # one eager pass first (kernel attributes, allocator warm-up); it must not move the BatchNorm buffers
            buffers = [b for root in (self.encoder, self.decoder) for b in root.buffers()]
            saved = [b.clone() for b in buffers]
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.no_grad():
                with torch.cuda.stream(side):
                    body()
                torch.cuda.current_stream().wait_stream(side)
                for b, s in zip(buffers, saved):
                    b.copy_(s)
                from rcfd import _lib
                # the eager pass recorded the step's weight packs and built their batched form (rcfd_pack_batch): that one
                # launch stays OUTSIDE the graph, issued right after every optimiser step, where it overlaps the launch
                # latency of the next replay instead of delaying the first convolutions
                pack = self._cache.get(('pack_batch', self.compute_dtype)) if self.multistream and not self.x3 else None
                if pack is not None:
                    pack['table'].run()
                l0 = _lib.launch_count
                graph, graph_b = torch.cuda.CUDAGraph(), None
                self._external_pack = pack is not None
                try:
                    if split_off is None:
                        with torch.cuda.graph(graph, stream=engine.capture_stream(dev)):
                            loss, grads = body()
                    else:
                        with torch.cuda.stream(side):          # the split unpack tables are built by an eager two-part pass
                            body(0)
                            body(1)
                        torch.cuda.current_stream().wait_stream(side)
                        for b, s in zip(buffers, saved):
                            b.copy_(s)
                        l0 = _lib.launch_count
                        graph_b = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(graph, stream=engine.capture_stream(dev)):
                            body(0)
                        with torch.cuda.graph(graph_b, stream=engine.capture_stream(dev), pool=graph.pool()):
                            loss, grads = body(1)
                finally:
                    self._external_pack = False
                    self._grad_split_addr = None
                self.last_capture_launches = _lib.launch_count - l0 + 1 + (1 if pack is not None else 0)      # + Adam (+ pack)
            entry = {'graph': graph, 'graph_b': graph_b, 'split_off': split_off, 'static': static, 'loss': loss, 'grads': grads,
                     'pack': pack}
            self._train_graphs[key] = entry
        return entry

    def _deliver_grads(self, param_grads, hook=True):
        for p, g in param_grads:
            if g is p.grad:              # written in place (rcfd.optim.FusedAdam flat buffers)
                continue
            if getattr(p, '_rcfd_flat', False) and p.grad is not None:
                p.grad.copy_(g)          # flat-buffer mode overwrites
                continue
            if p.grad is None:
                p.grad = g
            else:
                p.grad.add_(g)
        if hook and self.grad_hook is not None:
            self.grad_hook(param_grads)

    # ------------------------------------------------------------------ loss
    def compute_loss(self, image, output_depth, ground_truth, lidar_map, loss_func, w_smoothness,
                     loss_smoothness_kernel_size, validity_map_loss_smoothness, w_lidar_loss):
        """Reference src/fusionnet_model.py:172-302.  The canonical training configuration
        (single scale, 'l1', w_lidar_loss > 0, w_smoothness == 0) is one fused, sync-free
        masked-L1 kernel; every other combination follows the reference formula with tensor ops."""
        single = not isinstance(output_depth, list)
        if (single and loss_func == 'l1' and w_lidar_loss > 0.0 and not w_smoothness > 0.0
                and output_depth.is_cuda and output_depth.shape == ground_truth.shape):
            loss = losses.MaskedL1.apply(output_depth, ground_truth, lidar_map, float(w_lidar_loss))
            return loss, {'loss': loss, 'loss_supervised': loss, 'loss_smoothness': 0.0, 'loss_lidar': 0.0}

        loss_supervised, loss_smoothness, loss_lidar = 0.0, 0.0, 0.0
        if w_lidar_loss > 0.0:
            ground_truth = ground_truth * (lidar_map <= 0.0).to(ground_truth.dtype)
        valid_gt = ground_truth > 0
        valid_lidar = lidar_map > 0
        outputs = output_depth if isinstance(output_depth, list) else [output_depth]
        pick = {'l1': losses.l1_loss, 'l2': losses.l2_loss, 'smoothl1': losses.smooth_l1_loss}
        if loss_func not in pick:
            raise ValueError('No such loss: {}'.format(loss_func))
        for scale, output in enumerate(outputs):
            th, tw = ground_truth.shape[-2:]
            if output.shape[-2] > th and output.shape[-1] > tw:
                output = torch.nn.functional.interpolate(output, size=(th, tw), mode='bilinear', align_corners=True)
            w_scale = 1.0 / (2 ** (len(outputs) - scale - 1))
            loss_supervised = loss_supervised + w_scale * pick[loss_func](output[valid_gt], ground_truth[valid_gt])
            if w_lidar_loss > 0.0:
                loss_lidar = loss_lidar + w_scale * pick[loss_func](output[valid_lidar], lidar_map[valid_lidar])
            if w_smoothness > 0.0:
                if loss_smoothness_kernel_size <= 1:
                    loss_smoothness = loss_smoothness + w_scale * losses.smoothness_loss_func(image=image, predict=output)
                else:
                    fs = [1, 1, loss_smoothness_kernel_size, loss_smoothness_kernel_size]
                    loss_smoothness = loss_smoothness + w_scale * losses.sobel_smoothness_loss_func(
                        image=image, predict=output, weights=validity_map_loss_smoothness, filter_size=fs)
        loss = loss_supervised + w_smoothness * loss_smoothness + w_lidar_loss * loss_lidar
        return loss, {'loss': loss, 'loss_supervised': loss_supervised, 'loss_smoothness': loss_smoothness,
                      'loss_lidar': loss_lidar}

    # ------------------------------------------------------------------ state management
    def _state_tensors(self):
        if getattr(self, '_state_list', None) is None:
            self._state_list = [t for root in (self.encoder, self.decoder)
                                for t in list(root.parameters()) + list(root.buffers())]
        return self._state_list

    def parameters(self):
        """Encoder parameters then decoder parameters (reference :304-316; Adam state order)."""
        return list(self.encoder.parameters()) + list(self.decoder.parameters())

    def train(self):
        self.encoder.train()
        self.decoder.train()

    def eval(self):
        self.encoder.eval()
        self.decoder.eval()

    def to(self, device):
        self.device = device
        self.encoder.to(device)
        self.decoder.to(device)
        self._invalidate()

    def _modules_bare(self):
        enc = self.encoder.module if isinstance(self.encoder, torch.nn.DataParallel) else self.encoder
        dec = self.decoder.module if isinstance(self.decoder, torch.nn.DataParallel) else self.decoder
        return enc, dec

    # The reference always wraps encoder / decoder in torch.nn.DataParallel before it saves or restores
    # (src/fusionnet_main.py:198, 727-731), so its checkpoints carry 'module.'-prefixed keys and its strict
    # load_state_dict expects them.  True: write that format (loadable by the reference's run / train scripts);
    # restore_model accepts both.
    REFERENCE_CHECKPOINT_KEYS = True

    def save_model(self, checkpoint_path, step, optimizer):
        """Same checkpoint dictionary as the reference (:347-368), 'module.'-prefixed keys like the files the reference
        writes (see REFERENCE_CHECKPOINT_KEYS); the optimizer state is torch.optim.Adam's layout."""
        pre = 'module.' if self.REFERENCE_CHECKPOINT_KEYS else ''
        enc, dec = self._modules_bare()
        torch.save({'train_step': step,
                    'optimizer_state_dict': optimizer.state_dict(),
                    'encoder_state_dict': {pre + k: v for k, v in enc.state_dict().items()},
                    'decoder_state_dict': {pre + k: v for k, v in dec.state_dict().items()}}, checkpoint_path)

    @staticmethod
    def _strip_module_prefix(state):
        # the reference saves after data_parallel(), so its keys carry a 'module.' prefix
        return {(k[len('module.'):] if k.startswith('module.') else k): v for k, v in state.items()}

    def restore_model(self, checkpoint_path, optimizer=None):
        """Accepts both bare and 'module.'-prefixed (reference DataParallel) keys (:370-393)."""
        checkpoint = torch.load(checkpoint_path, map_location=self.device, weights_only=False)
        self.encoder.load_state_dict(self._strip_module_prefix(checkpoint['encoder_state_dict']))
        self.decoder.load_state_dict(self._strip_module_prefix(checkpoint['decoder_state_dict']))
        self._invalidate()
        engine.note_params_changed()
        if optimizer is not None:
            optimizer.load_state_dict(checkpoint['optimizer_state_dict'])
        return checkpoint['train_step'], optimizer

    def data_parallel(self):
        """The reference wraps encoder / decoder in torch.nn.DataParallel (:395-401).  Here
        multi-GPU is one process per GPU (torchrun) with an NCCL gradient all-reduce:
        see rcfd.parallel.DistributedGradSync.  Single process: nothing to do."""
        from rcfd import parallel
        parallel.attach_if_distributed(self)

    def log_summary(self, *args, **kwargs):
        """TensorBoard summaries are observability, out of scope for the hot path."""
        return None
