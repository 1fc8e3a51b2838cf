"""
Network graphs with the reference's class names, constructor arguments, attribute names and
state_dict keys (reference: src/networks.py).  The classes hold parameters and describe the
topology; rcfd.engine executes it on librcfd_b200.so (NHWC, fused epilogues, no materialised
up-sample / concat).  Supported = what the shipped configs reach:
FusionNetEncoder(fusion_type='weight_and_project'), ResNetEncoder, FullyConnectedEncoder,
RadarNetV1Encoder, MultiScaleDecoder(n_resolution=1, deconv_type='up').
"""
import torch

import net_utils

_BLOCKS_PER_STAGE = {18: [2, 2, 2, 2], 34: [3, 4, 6, 3]}


def _stage_blocks(n_layer, depth):
    if n_layer not in _BLOCKS_PER_STAGE:
        raise ValueError('Only supports 18, 34 layer architecture')
    n_blocks = list(_BLOCKS_PER_STAGE[n_layer])
    while len(n_blocks) < depth - 1:           # deeper pyramids repeat the last stage
        n_blocks.append(n_blocks[-1])
    assert depth < 8, 'Does not support network depth of 8 or more'
    assert depth == len(n_blocks) + 1
    return n_blocks


def _stage(n_block, in_channels, out_channels, stride, weight_initializer, activation, use_batch_norm):
    blocks = []
    for n in range(n_block):
        blocks.append(net_utils.ResNetBlock(in_channels if n == 0 else out_channels, out_channels,
                                            stride if n == 0 else 1, weight_initializer, activation,
                                            use_batch_norm))
    return torch.nn.Sequential(*blocks)


class ResNetEncoder(torch.nn.Module):
    """ResNet encoder with skip connections (reference src/networks.py:8-268)."""

    def __init__(self, n_layer, input_channels=3, n_filters=[32, 64, 128, 256, 256],
                 weight_initializer='kaiming_uniform', activation_func='leaky_relu', use_batch_norm=False):
        super(ResNetEncoder, self).__init__()
        n_blocks = _stage_blocks(n_layer, len(n_filters))
        act = net_utils.activation_func(activation_func)
        self.n_filters = list(n_filters)
        self.conv1 = net_utils.Conv2d(input_channels, n_filters[0], 7, 2, weight_initializer, act, use_batch_norm)
        self.max_pool = torch.nn.MaxPool2d(kernel_size=3, stride=2, padding=1)    # descriptor only
        for level in range(2, 8):
            idx = level - 1
            if idx < len(n_filters):
                stage = _stage(n_blocks[idx - 1], n_filters[idx - 1], n_filters[idx], 1 if level == 2 else 2,
                               weight_initializer, act, use_batch_norm)
            else:
                stage = None
            setattr(self, 'blocks%d' % level, stage)

    def forward(self, x):
        from rcfd import engine
        return engine.standalone(self, 'resnet_encoder', x)


class FusionNetEncoder(torch.nn.Module):
    """Two-branch (image / depth+response) ResNet encoder with gated per-level fusion
    ``sigmoid(BN(Ww d)) * BN(Wp d) + image`` (reference src/networks.py:270-1005)."""

    def __init__(self, n_layer=18, input_channels_image=3, input_channels_depth=3,
                 n_filters_encoder_image=[32, 64, 128, 256, 256], n_filters_encoder_depth=[32, 64, 128, 256, 256],
                 weight_initializer='kaiming_uniform', activation_func='leaky_relu', use_batch_norm=False,
                 fusion_type='add'):
        super(FusionNetEncoder, self).__init__()
        if fusion_type != 'weight_and_project':
            raise ValueError("Unsupported fusion type on the B200 path: {} (every shipped config uses "
                             "'weight_and_project')".format(fusion_type))
        assert len(n_filters_encoder_image) == len(n_filters_encoder_depth)
        self.fusion_type = fusion_type
        fi, fd = list(n_filters_encoder_image), list(n_filters_encoder_depth)
        self.n_filters_image, self.n_filters_depth = fi, fd
        n_blocks = _stage_blocks(n_layer, len(fi))
        act = net_utils.activation_func(activation_func)
        sig = net_utils.activation_func('sigmoid')
        lin = net_utils.activation_func('linear')

        def add_fusion(level, idx):
            setattr(self, 'conv%d_weight' % level,
                    net_utils.Conv2d(fd[idx], fi[idx], 1, 1, weight_initializer, sig, use_batch_norm))
            setattr(self, 'conv%d_project' % level,
                    net_utils.Conv2d(fd[idx], fi[idx], 1, 1, weight_initializer, lin, use_batch_norm))

        self.conv1_image = net_utils.Conv2d(input_channels_image, fi[0], 7, 2, weight_initializer, act, use_batch_norm)
        self.conv1_depth = net_utils.Conv2d(input_channels_depth, fd[0], 7, 2, weight_initializer, act, use_batch_norm)
        add_fusion(1, 0)
        self.max_pool = torch.nn.MaxPool2d(kernel_size=3, stride=2, padding=1)    # descriptor only
        for level in range(2, 8):
            idx = level - 1
            if idx < len(fi):
                stride = 1 if level == 2 else 2
                setattr(self, 'blocks%d_image' % level,
                        _stage(n_blocks[idx - 1], fi[idx - 1], fi[idx], stride, weight_initializer, act, use_batch_norm))
                setattr(self, 'blocks%d_depth' % level,
                        _stage(n_blocks[idx - 1], fd[idx - 1], fd[idx], stride, weight_initializer, act, use_batch_norm))
                add_fusion(level, idx)
            else:
                for name in ('blocks%d_image', 'blocks%d_depth', 'conv%d_weight', 'conv%d_project'):
                    setattr(self, name % level, None)

    def forward(self, image, depth):
        from rcfd import engine
        return engine.standalone(self, 'fusionnet_encoder', image, depth=depth)


class FullyConnectedEncoder(torch.nn.Module):
    """Radar point MLP: 6 x (Linear + LeakyReLU) (reference src/networks.py:1007-1067)."""

    def __init__(self, input_channels=3, n_neurons=[32, 64, 96, 128, 256], latent_size=29 * 10,
                 weight_initializer='kaiming_uniform', activation_func='leaky_relu'):
        super(FullyConnectedEncoder, self).__init__()
        act = net_utils.activation_func(activation_func)
        sizes = [input_channels] + list(n_neurons[:5]) + [latent_size]
        self.mlp = torch.nn.Sequential(*[
            net_utils.FullyConnected(sizes[i], sizes[i + 1], weight_initializer, act) for i in range(6)])

    def forward(self, x):
        return self.mlp(x)


class RadarNetV1Encoder(torch.nn.Module):
    """Image ResNet encoder + per-point column ROI pooling + point MLP
    (reference src/networks.py:1151-1256)."""

    def __init__(self, input_channels_image=3, input_channels_depth=3, input_patch_size_image=(900, 288),
                 n_filters_encoder_image=[32, 64, 128, 128, 128], n_neurons_encoder_depth=[32, 64, 128, 128, 128],
                 latent_size_depth=128 * 29 * 10, weight_initializer='kaiming_uniform', activation_func='leaky_relu',
                 use_batch_norm=False):
        super(RadarNetV1Encoder, self).__init__()
        self.n_neuron_latent_depth = n_neurons_encoder_depth[-1]
        self.encoder_image = ResNetEncoder(18, input_channels_image, n_filters_encoder_image, weight_initializer,
                                           activation_func, use_batch_norm)
        self.encoder_depth = FullyConnectedEncoder(input_channels_depth, n_neurons_encoder_depth, latent_size_depth,
                                                   weight_initializer, activation_func)
        self.input_patch_size_image = input_patch_size_image

    def forward(self, image, points, b_boxes):
        from rcfd import engine
        return engine.standalone(self, 'radarnet_encoder', image, points=points, boxes=b_boxes)


class MultiScaleDecoder(torch.nn.Module):
    """U-Net decoder (reference src/networks.py:1337-1657).  n_resolution > 1: output{b} convs after deconv{b} (b = 1..3,
    present when n_resolution > b) whose logits, up-sampled 2x bilinearly, join the next block's skip (one more skip
    channel for deconv{b-1}); modules are registered in the reference's order (deconv3, output3, deconv2, output2, ...),
    which is the order of parameters() and of the optimiser state."""

    def __init__(self, input_channels=256, output_channels=1, n_resolution=1, n_filters=[256, 128, 64, 32, 16],
                 n_skips=[256, 128, 64, 32, 0], weight_initializer='kaiming_uniform', activation_func='leaky_relu',
                 output_func='linear', use_batch_norm=False, deconv_type='up'):
        super(MultiScaleDecoder, self).__init__()
        depth = len(n_filters)
        assert depth < 8, 'Does not support network depth of 8 or more'
        assert n_resolution > 0 and n_resolution < depth
        if n_resolution > 4:
            raise ValueError('n_resolution must be 1 .. 4 (the reference builds output0 .. output3)')
        if output_func != 'linear':
            raise ValueError("output_func must be 'linear' (reference fusionnet_model.py:131, radarnet_model.py:94)")
        self.n_resolution = n_resolution
        self.output_func = output_func
        act = net_utils.activation_func(activation_func)
        self.n_blocks = depth
        in_channels = input_channels
        for i in range(depth):
            b = depth - 1 - i                   # deconv{depth-1} ... deconv0
            extra = output_channels if (b <= 2 and n_resolution > b + 1) else 0      # the up-sampled logits of output{b+1}
            setattr(self, 'deconv%d' % b,
                    net_utils.DecoderBlock(in_channels, n_skips[i] + extra, n_filters[i], weight_initializer, act,
                                           use_batch_norm, deconv_type))
            in_channels = n_filters[i]
            if 1 <= b <= 3 and n_resolution > b:
                setattr(self, 'output%d' % b, net_utils.Conv2d(in_channels, output_channels, 3, 1, weight_initializer,
                                                               net_utils.activation_func(output_func), False))
        for b in range(depth, 7):
            setattr(self, 'deconv%d' % b, None)
        self.output0 = net_utils.Conv2d(in_channels, output_channels, 3, 1, weight_initializer,
                                        net_utils.activation_func(output_func), False)

    def forward(self, x, skips, shape=None):
        from rcfd import engine
        return engine.standalone(self, 'decoder', x, skips=skips, shape=shape)
