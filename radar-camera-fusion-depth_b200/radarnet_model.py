"""
RadarNetModel (stage 1: which pixels does a radar return belong to) with the reference's
constructor / method surface (reference: src/radarnet_model.py:7-258), forward executed on
librcfd_b200.so: image ResNet encoder once per image, column ROI pooling + point MLP +
U-Net decoder once per radar point.  Forward (inference) is the accelerated path of this
round; the stage-1 training loop is listed as "next" in SURVEY.md 8f.
"""
import torch

import networks
from rcfd import engine, ops


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

    def forward(self, image, point, bounding_boxes, return_logits=True):
        """image N x 3 x H x W (already edge-padded), point sum(K_i) x 3, bounding_boxes list of K_i x 4
        -> sum(K_i) x 1 x ph x pw logits or sigmoid responses (reference :102-124)."""
        if self.encoder.training and torch.is_grad_enabled():
            raise NotImplementedError('RadarNet training (backward) is not part of this round; call under '
                                      'torch.no_grad() / model.eval() for stage-1 inference')
        ctx = engine.Context(self.compute_dtype, False, image.device, cache=self._cache, engine=self.conv_engine, x3=self.x3)
        img, s2d = engine.stem_input(ctx, image)
        latent, skips = engine.radarnet_encoder(ctx, self.encoder, img, point, bounding_boxes, stem_s2d=s2d)
        dec = self.decoder
        out0 = dec.output0
        # logits or sigmoid straight from output0's epilogue
        saved = out0.act_kind
        try:
            out0.act_kind = 'linear' if return_logits else 'sigmoid'
            x = latent
            n = len(skips) - 1
            for b in range(dec.n_blocks - 1, -1, -1):
                blk = getattr(dec, 'deconv%d' % b)
                if n >= 0:
                    x = engine.decoder_block(ctx, blk, x, skips[n], None)
                    n -= 1
                else:
                    x = engine.decoder_block(ctx, blk, x, None, self.input_patch_size_image)
            out = ops.conv2d(x, ctx.weight(out0), 1, 3, 1, act=engine._ACT[out0.act_kind], out_f32=True,
                             engine=ctx.engine)
        finally:
            out0.act_kind = saved
        k, h, w, _ = out.shape
        return out.view(k, 1, h, w)

    def compute_loss(self, logits, ground_truth, validity_map, w_positive_class=1.0):
        """Weighted BCE over valid pixels (reference :126-167); tensor-op formula ("next")."""
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
        """Stage-1 inference shards by image: replicas only, nothing to wrap."""
        return None

    def log_summary(self, *args, **kwargs):
        return None
