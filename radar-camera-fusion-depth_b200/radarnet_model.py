"""
RadarNetModel (stage 1: which pixels does a radar return belong to) with the reference's
constructor / method surface (reference: src/radarnet_model.py:7-258), executed on
librcfd_b200.so: image ResNet encoder once per image, column ROI pooling + point MLP +
U-Net decoder once per radar point.  Inference (forward + S2 scatter, radarnet_main.forward)
and the training step (forward -> weighted BCE -> backward, radarnet_main.train) both run
through the C-ABI; ``forward`` in train mode returns logits that support ``loss.backward()``
(one autograd node whose backward replays the engine's tape, like FusionNetModel).
"""
import torch

import networks
from rcfd import engine, ops


class _RadarNetFunction(torch.autograd.Function):
    """One autograd node for encoder + decoder (training)."""

    @staticmethod
    def forward(fctx, model, image, point, boxes, *params):
        out, ectx = model._run(image, point, boxes, True, record=True)
        fctx.model, fctx.ectx, fctx.out = model, ectx, out
        fctx.set_materialize_grads(False)
        k, h, w, _ = out.shape
        return out.view(k, 1, h, w)

    @staticmethod
    def backward(fctx, grad_out):
        ectx = fctx.ectx
        n_in = 4 + len(fctx.model.parameters())
        if grad_out is None or ectx is None or ectx.tape is None:
            return (None,) * n_in
        tape = ectx.tape
        tape.set_grad(fctx.out, grad_out.contiguous().float().view(fctx.out.shape))
        tape.backward()
        for p, g in tape.param_grads:
            if g is p.grad:
                continue
            if getattr(p, '_rcfd_flat', False) and p.grad is not None:
                p.grad.copy_(g)
            elif p.grad is None:
                p.grad = g
            else:
                p.grad.add_(g)
        if fctx.model.grad_hook is not None:
            fctx.model.grad_hook(tape.param_grads)
        tape.param_grads = []
        fctx.ectx = None
        return (None,) * n_in


class WeightedBCE(torch.autograd.Function):
    """sum(v * BCEWithLogits(x, t, pos_weight)) / sum(v)  (reference src/radarnet_model.py:148-161) as one fused,
    synchronisation-free kernel pair (rcfd_bce_logits_loss)."""

    @staticmethod
    def forward(ctx, logits, ground_truth, validity_map, w_positive_class):
        loss, dl = ops.bce_logits_loss(logits.detach().float(), ground_truth.float(), validity_map.float(),
                                       float(w_positive_class), want_grad=True)
        ctx.save_for_backward(dl)
        return loss.view(())

    @staticmethod
    def backward(ctx, g):
        (dl,) = ctx.saved_tensors
        return dl * g, None, None, None


class RadarNetModel(object):

    def __init__(self, input_channels_image, input_channels_depth, input_patch_size_image, encoder_type,
                 n_filters_encoder_image, n_neurons_encoder_depth, decoder_type, n_filters_decoder,
                 weight_initializer='kaiming_uniform', activation_func='leaky_relu', device=torch.device('cuda')):
        self.input_patch_size_image = input_patch_size_image
        self.device = device
        self.compute_dtype = torch.float32
        self.x3 = 0
        self.precision = 'fp32'
        self.conv_engine = ops.ENGINE_AUTO
        self._cache = {}
        self.grad_hook = None          # called with the list of (param, grad) after backward (data parallel)
        height, width = input_patch_size_image
        latent_size_depth = int(height // 32.0) * int(width // 32.0) * n_neurons_encoder_depth[-1]
        if 'radarnetv1' in encoder_type:
            self.encoder = networks.RadarNetV1Encoder(
                input_channels_image=input_channels_image, input_channels_depth=input_channels_depth,
                input_patch_size_image=input_patch_size_image, n_filters_encoder_image=n_filters_encoder_image,
                n_neurons_encoder_depth=n_neurons_encoder_depth, latent_size_depth=latent_size_depth,
                weight_initializer=weight_initializer, activation_func=activation_func,
                use_batch_norm='batch_norm' in encoder_type)
        else:
            raise ValueError('Encoder type {} not supported.'.format(encoder_type))
        n_skips = list(n_filters_encoder_image[:-1])[::-1] + [0]
        latent_channels = n_filters_encoder_image[-1] + n_neurons_encoder_depth[-1]
        if 'multiscale' in decoder_type:
            self.decoder = networks.MultiScaleDecoder(
                input_channels=latent_channels, output_channels=1, n_resolution=1, n_filters=n_filters_decoder,
                n_skips=n_skips, weight_initializer=weight_initializer, activation_func=activation_func,
                output_func='linear', use_batch_norm='batch_norm' in decoder_type, deconv_type='up')
        else:
            raise ValueError('Decoder type {} not supported.'.format(decoder_type))
        self.to(self.device)

    def set_precision(self, precision):
        self.compute_dtype = {'fp32': torch.float32, 'bf16': torch.bfloat16, 'bf16x3': torch.float32, 'bf16x6': torch.float32}[precision]
        self.x3 = {'bf16x3': 2, 'bf16x6': 3}.get(precision, 0)       # tensor-core parity mode (see FusionNetModel.set_precision)
        self.precision = precision
        self._cache.clear()
        return self

    def _run(self, image, point, bounding_boxes, return_logits, record=False):
        training = self.encoder.training
        with ops.hold_allocations():
            ctx = engine.Context(self.compute_dtype, training, image.device, cache=self._cache, record=record,
                                 engine=self.conv_engine, x3=self.x3)
            img, s2d = engine.stem_input(ctx, image)
            latent, skips = engine.radarnet_encoder(ctx, self.encoder, img, point, bounding_boxes, stem_s2d=s2d)
            dec = self.decoder
            x = latent
            n = len(skips) - 1
            for b in range(dec.n_blocks - 1, -1, -1):
                blk = getattr(dec, 'deconv%d' % b)
                if n >= 0:
                    x = engine.decoder_block(ctx, blk, x, skips[n], None)
                    n -= 1
                else:
                    x = engine.decoder_block(ctx, blk, x, None, self.input_patch_size_image)
            # logits or sigmoid straight from output0's epilogue (float32 out)
            act = ops.ACT_NONE if return_logits else ops.ACT_SIGMOID
            out0 = dec.output0
            out = ops.conv2d(x, ctx.weight(out0), 1, 3, 1, act=act, out_f32=True, engine=ctx.engine)
            if ctx.tape is not None:
                assert return_logits
                k, h, w, _ = out.shape
                cpad = engine.CPAD_DY if (ctx.dtype == torch.bfloat16 or ctx.x3) else 8
                engine._record_conv_backward(
                    ctx, out0, x, None, None, out, None, True,
                    pre=lambda dd: ops.nchw_to_nhwc(dd.reshape(k, 1, h, w).float(), ctx.dtype, cpad=cpad))
            engine.finish_bn_counters(ctx)
        return out, ctx

    def forward(self, image, point, bounding_boxes, return_logits=True):
        """image N x 3 x H x W (already edge-padded), point sum(K_i) x 3, bounding_boxes list of K_i x 4
        -> sum(K_i) x 1 x ph x pw logits or sigmoid responses (reference :102-124).  In train mode with grad enabled the
        returned logits carry the backward of the whole column."""
        params = self.parameters()
        if self.encoder.training and torch.is_grad_enabled() and any(p.requires_grad for p in params):
            if not return_logits:
                raise ValueError('training forward returns logits (the loss is BCE with logits, reference :148-153)')
            return _RadarNetFunction.apply(self, image, point, bounding_boxes, *params)
        out, _ = self._run(image, point, bounding_boxes, return_logits)
        k, h, w, _ = out.shape
        return out.view(k, 1, h, w)

    def compute_loss(self, logits, ground_truth, validity_map, w_positive_class=1.0):
        """Weighted BCE over valid pixels (reference :126-167): one fused kernel pair on CUDA tensors, the reference's
        tensor-op formula otherwise."""
        if logits.is_cuda:
            loss = WeightedBCE.apply(logits, ground_truth, validity_map, float(w_positive_class))
            return loss, {'loss': loss}
        pw = torch.tensor(w_positive_class, device=logits.device)
        loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, ground_truth, reduction='none',
                                                                    pos_weight=pw)
        loss = torch.sum(validity_map * loss) / torch.sum(validity_map)
        return loss, {'loss': loss}

    def parameters(self):
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
        self._cache.clear()

    def save_model(self, checkpoint_path, step, optimizer):
        """Reference key names (:212-233); 'module.'-prefixed like the files the reference writes after data_parallel()
        (src/radarnet_main.py:203), restore_model accepts both."""
        torch.save({'train_step': step,
                    'radarnet_optimizer_state_dict': optimizer.state_dict(),
                    'radarnet_encoder_state_dict': {'module.' + k: v for k, v in self.encoder.state_dict().items()},
                    'radarnet_decoder_state_dict': {'module.' + k: v for k, v in self.decoder.state_dict().items()}},
                   checkpoint_path)

    def restore_model(self, checkpoint_path, optimizer=None):
        strip = lambda sd: {(k[7:] if k.startswith('module.') else k): v for k, v in sd.items()}
        checkpoint = torch.load(checkpoint_path, map_location=self.device, weights_only=False)
        self.encoder.load_state_dict(strip(checkpoint['radarnet_encoder_state_dict']))
        self.decoder.load_state_dict(strip(checkpoint['radarnet_decoder_state_dict']))
        self._cache.clear()
        if optimizer is not None:
            optimizer.load_state_dict(checkpoint['radarnet_optimizer_state_dict'])
        return checkpoint['train_step'], optimizer

    def data_parallel(self):
        """Inference shards by image (replicas only).  Training: one process per GPU, NCCL gradient all-reduce
        (rcfd.parallel.DistributedGradSync), attached when torch.distributed is initialised."""
        from rcfd import parallel
        parallel.attach_if_distributed(self)

    def log_summary(self, *args, **kwargs):
        return None
