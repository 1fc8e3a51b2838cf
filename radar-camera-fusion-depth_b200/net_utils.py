"""
Layer / block library with the reference's names and constructor signatures
(reference: src/net_utils.py).  The modules here are PARAMETER CONTAINERS with the
reference's exact state_dict layout (``conv.weight`` OIHW float32, ``batch_norm.*``);
all arithmetic runs in librcfd_b200.so through rcfd.engine -- there is no PyTorch
compute path and no CPU fallback.  Calling a block on its own (NCHW float CUDA tensor)
runs the same kernels in inference mode.
"""
import torch


class Activation(torch.nn.Module):
    """What the reference's activation_func() hands to its layers (src/net_utils.py:4-23);
    here only a tag the engine turns into a fused epilogue."""

    def __init__(self, kind):
        super(Activation, self).__init__()
        self.kind = kind

    def forward(self, x):
        from rcfd import engine
        return engine.standalone_activation(x, self.kind)

    def extra_repr(self):
        return self.kind


def activation_func(activation_fn):
    """Select activation function (reference: src/net_utils.py:4-23; leaky slope 0.20)."""
    if 'linear' in activation_fn:
        return None
    elif 'leaky_relu' in activation_fn:
        return Activation('leaky_relu')
    elif 'sigmoid' in activation_fn:
        return Activation('sigmoid')
    elif 'relu' in activation_fn or 'elu' in activation_fn:
        raise ValueError('Unsupported activation function on the B200 path: {} '
                         '(the shipped configs use leaky_relu / sigmoid / linear)'.format(activation_fn))
    else:
        raise ValueError('Unsupported activation function: {}'.format(activation_fn))


def _act_kind(activation):
    if activation is None:
        return 'linear'
    if isinstance(activation, Activation):
        return activation.kind
    if isinstance(activation, torch.nn.LeakyReLU):
        if abs(activation.negative_slope - 0.2) > 1e-12:
            raise ValueError('only LeakyReLU(0.2) (reference activation_func) is supported')
        return 'leaky_relu'
    if isinstance(activation, torch.nn.Sigmoid):
        return 'sigmoid'
    raise ValueError('Unsupported activation module: {}'.format(activation))


def _init_weight(weight, weight_initializer):
    # reference src/net_utils.py:71-77: 'kaiming_uniform' matches no branch, so torch's
    # default init (kaiming_uniform_(a=sqrt(5))) stays.
    if weight_initializer == 'kaiming_normal':
        torch.nn.init.kaiming_normal_(weight)
    elif weight_initializer == 'xavier_normal':
        torch.nn.init.xavier_normal_(weight)
    elif weight_initializer == 'xavier_uniform':
        torch.nn.init.xavier_uniform_(weight)


class Conv2d(torch.nn.Module):
    """conv(bias=False, pad=k//2) -> BatchNorm? -> activation?  (reference src/net_utils.py:29-91)"""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1,
                 weight_initializer='kaiming_uniform', activation_func=Activation('leaky_relu'),
                 use_batch_norm=False):
        super(Conv2d, self).__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride = kernel_size, stride
        self.use_batch_norm = use_batch_norm
        self.conv = torch.nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride,
                                    padding=kernel_size // 2, bias=False)
        _init_weight(self.conv.weight, weight_initializer)
        self.act_kind = _act_kind(activation_func)
        if use_batch_norm:
            self.batch_norm = torch.nn.BatchNorm2d(out_channels)

    def forward(self, x):
        from rcfd import engine
        return engine.standalone(self, 'conv', x)


class UpConv2d(torch.nn.Module):
    """nearest interpolate(size=shape) -> Conv2d  (reference src/net_utils.py:156-198);
    the up-sampled tensor is never materialised."""

    def __init__(self, in_channels, out_channels, kernel_size=3, weight_initializer='kaiming_uniform',
                 activation_func=Activation('leaky_relu'), use_batch_norm=False):
        super(UpConv2d, self).__init__()
        self.conv = Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=1,
                           weight_initializer=weight_initializer, activation_func=activation_func,
                           use_batch_norm=use_batch_norm)

    def forward(self, x, shape):
        from rcfd import engine
        return engine.standalone(self, 'upconv', x, shape=shape)


class FullyConnected(torch.nn.Module):
    """Linear(bias) -> activation (reference src/net_utils.py:201-247); dropout_rate must be 0."""

    def __init__(self, in_features, out_features, weight_initializer='kaiming_uniform',
                 activation_func=Activation('leaky_relu'), dropout_rate=0.00):
        super(FullyConnected, self).__init__()
        if dropout_rate > 0.0:
            raise ValueError('dropout is not used by any shipped config and is not implemented')
        self.fully_connected = torch.nn.Linear(in_features, out_features)
        _init_weight(self.fully_connected.weight, weight_initializer)
        self.act_kind = _act_kind(activation_func)
        if self.act_kind != 'leaky_relu':
            raise ValueError('FullyConnected supports leaky_relu only (reference MLP, src/networks.py:1033-1063)')

    def forward(self, x):
        from rcfd import ops
        return ops.linear_leaky(x, self.fully_connected.weight, self.fully_connected.bias)


class ResNetBlock(torch.nn.Module):
    """Basic ResNet block (reference src/net_utils.py:253-323): the activation is applied to
    conv2 before the residual add and again after it; the projection has no BN / activation
    and exists in every block but runs only when shape or channels change."""

    def __init__(self, in_channels, out_channels, stride=1, weight_initializer='kaiming_uniform',
                 activation_func=Activation('leaky_relu'), use_batch_norm=False):
        super(ResNetBlock, self).__init__()
        if _act_kind(activation_func) != 'leaky_relu':
            raise ValueError('ResNetBlock supports leaky_relu only')
        self.in_channels, self.out_channels, self.stride = in_channels, out_channels, stride
        self.conv1 = Conv2d(in_channels, out_channels, 3, stride, weight_initializer, activation_func, use_batch_norm)
        self.conv2 = Conv2d(out_channels, out_channels, 3, 1, weight_initializer, activation_func, use_batch_norm)
        self.projection = Conv2d(in_channels, out_channels, 1, stride, weight_initializer, None, False)

    def forward(self, x):
        from rcfd import engine
        return engine.standalone(self, 'resblock', x)


class DecoderBlock(torch.nn.Module):
    """up-conv -> cat skip -> conv (reference src/net_utils.py:473-569, deconv_type 'up');
    neither the up-sampled tensor nor the concat is materialised."""

    def __init__(self, in_channels, skip_channels, out_channels, weight_initializer='kaiming_uniform',
                 activation_func=Activation('leaky_relu'), use_batch_norm=False, deconv_type='up'):
        super(DecoderBlock, self).__init__()
        if deconv_type != 'up':
            raise ValueError("deconv_type '{}' is not supported: the reference hard-codes 'up' "
                             "(src/fusionnet_main.py:190,718)".format(deconv_type))
        self.skip_channels = skip_channels
        self.deconv_type = deconv_type
        self.deconv = UpConv2d(in_channels, out_channels, 3, weight_initializer, activation_func, use_batch_norm)
        self.conv = Conv2d(skip_channels + out_channels, out_channels, 3, 1, weight_initializer, activation_func,
                           use_batch_norm)

    def forward(self, x, skip=None, shape=None):
        from rcfd import engine
        return engine.standalone(self, 'decoder_block', x, skip=skip, shape=shape)


class OutlierRemoval(object):
    """Local-minimum outlier filter for sparse depth (reference src/net_utils.py:575-638),
    one fused min-filter kernel instead of pad + max_pool + where."""

    def __init__(self, kernel_size=7, threshold=1.5):
        self.kernel_size = kernel_size
        self.threshold = threshold

    def remove_outliers(self, depth):
        from rcfd import ops
        return ops.outlier_removal(depth, self.kernel_size, self.threshold)
