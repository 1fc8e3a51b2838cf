"""Shared helpers for the parity tests (CPU side only uses torch + numpy)."""
import os
from collections import OrderedDict

import numpy as np
import torch

from rcfd import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)


def relerr(a, b):
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def fusionnet_state_template(cfg):
    """An ordered (name -> zero tensor) dict with exactly the reference's state_dict
    keys, shapes and ORDER for a FusionNet config, built from the product's own module
    tree (CPU tensors; no CUDA needed).  rcfd.synth.fill_state_dict_ walks it in order,
    so the order must equal the reference's (checked against golden weights)."""
    import networks
    enc = networks.FusionNetEncoder(
        n_layer=18, input_channels_image=cfg['input_channels_image'],
        input_channels_depth=cfg['input_channels_depth'],
        n_filters_encoder_image=cfg['n_filters_encoder_image'],
        n_filters_encoder_depth=cfg['n_filters_encoder_depth'],
        weight_initializer=cfg['weight_initializer'], activation_func=cfg['activation_func'],
        use_batch_norm=True, fusion_type=cfg['fusion_type'])
    n_skips = cfg['n_filters_encoder_image'][:-1][::-1] + [0]
    dec = networks.MultiScaleDecoder(
        input_channels=cfg['n_filters_encoder_image'][-1], output_channels=1, n_resolution=cfg.get('n_resolution_decoder', 1),
        n_filters=cfg['n_filters_decoder'], n_skips=n_skips, weight_initializer=cfg['weight_initializer'],
        activation_func=cfg['activation_func'], output_func='linear', use_batch_norm=True, deconv_type='up')
    p = OrderedDict()
    for k, v in enc.state_dict().items():
        p['encoder.' + k] = v.detach().clone()
    for k, v in dec.state_dict().items():
        p['decoder.' + k] = v.detach().clone()
    return p


def synth_fusionnet_state(cfg, seed):
    p = fusionnet_state_template(cfg)
    synth.fill_state_dict_(p, seed)
    return p


def multires_case_inputs():
    """(golden, cfg, state dict, image, depth, per-scale loss weights) of tests/golden/fusionnet_multires3_2x64x96.npz: the
    weights are regenerated from the seed (rcfd.synth.fill_state_dict_), like the canonical-width fixtures."""
    g = load_golden('fusionnet_multires3_2x64x96')
    n, h, w, seed, nres = [int(v) for v in g['meta']]
    cfg = dict(synth.SMALL_FUSIONNET, n_resolution_decoder=nres)
    image, depth = synth.fusionnet_inputs(n, h, w, seed, 'quasi_dense')
    return g, cfg, seed, image, depth, [float(i + 1) for i in range(nres)]


def check_multires_against_golden(g, outs, named, loss, tol_out=1e-4, tol_grad=1e-3):
    """outputs at every scale, the multi-scale loss and every parameter gradient fingerprint written by the reference."""
    assert len(outs) == int(g['meta'][4])
    for i, o in enumerate(outs):
        assert relerr(o.detach().float().cpu(), g['depth%d' % i]) < tol_out, i
    assert abs(float(loss) - float(g['loss'])) < tol_out * abs(float(g['loss']))
    for i, k in enumerate(g['grad_names']):
        k = str(k)
        if g['grad_none'][i]:
            assert named[k].grad is None, k
        else:
            assert named[k].grad is not None, k
            assert abs(float(named[k].grad.double().sum()) - g['grad_sum'][i]) <= tol_grad * g['grad_abs'][i] + 1e-9, k
