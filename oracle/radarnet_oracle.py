"""
ORACLE (test infrastructure, NOT product code) -- CPU fp32 restatement of the RadarNet
stage-1 column.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg may import this file.

Parity status: PINNED (tests/golden/make_golden.py runs the unmodified reference
RadarNetModel.forward incl. torchvision.ops.roi_pool and stores its outputs).

roi_pool itself lives in torchvision (0.11.3 pinned by the reference's
requirements.txt:142, 0.26.0 in this image; not under /root/reference).  Its published
algorithm (torchvision/csrc/ops/cpu/roi_pool_kernel.cpp) is restated in ``roi_pool``
below and anchored on the reference's call sites src/networks.py:1232-1247.
"""
import math
import numpy as np
import torch
import torch.nn.functional as F

from fusionnet_oracle import conv_block, resnet_stage, multiscale_decoder, LEAKY_SLOPE


def _round_half_away(v):
    # C round(): half away from zero, evaluated in float32 like the kernel's T=float
    v = np.float32(v)
    return int(math.floor(float(v) + 0.5)) if v >= 0 else -int(math.floor(-float(v) + 0.5))


def roi_pool(feat, boxes_list, spatial_scale, output_size):
    """torchvision.ops.roi_pool semantics.  feat: N x C x H x W; boxes_list: list of
    K_i x 4 (x1, y1, x2, y2).  Returns sum(K_i) x C x ph x pw."""
    n, c, h, w = feat.shape
    ph, pw = output_size
    outs = []
    for b, boxes in enumerate(boxes_list):
        for box in boxes:
            x1, y1, x2, y2 = [np.float32(v) * np.float32(spatial_scale) for v in box.tolist()]
            sw, sh, ew, eh = (_round_half_away(x1), _round_half_away(y1),
                              _round_half_away(x2), _round_half_away(y2))
            rw = max(ew - sw + 1, 1)
            rh = max(eh - sh + 1, 1)
            bh = np.float32(rh) / np.float32(ph)
            bw = np.float32(rw) / np.float32(pw)
            out = torch.zeros(c, ph, pw, dtype=feat.dtype)
            for i in range(ph):
                hs = min(max(int(math.floor(np.float32(i) * bh)) + sh, 0), h)
                he = min(max(int(math.ceil(np.float32(i + 1) * bh)) + sh, 0), h)
                for j in range(pw):
                    ws = min(max(int(math.floor(np.float32(j) * bw)) + sw, 0), w)
                    we = min(max(int(math.ceil(np.float32(j + 1) * bw)) + sw, 0), w)
                    if he <= hs or we <= ws:
                        continue
                    out[:, i, j] = feat[b, :, hs:he, ws:we].amax(dim=(1, 2))
            outs.append(out)
    return torch.stack(outs, dim=0)


def roi_pool_vectorised(feat, boxes_list, spatial_scale, output_size):
    """Same values as ``roi_pool`` above (checked against it in tests/test_oracle_golden.py), with the bin loops
    replaced by gathers: a bin is at most (mh x mw) pixels, so the max runs over mh * mw shifted gathers.  Used at the
    BASELINE configs[2] size, where the scalar loops take minutes."""
    n, c, h, w = feat.shape
    ph, pw = output_size
    outs = []
    f32 = np.float32
    for b, boxes in enumerate(boxes_list):
        for box in boxes:
            x1, y1, x2, y2 = [f32(v) * f32(spatial_scale) for v in box.tolist()]
            sw, sh, ew, eh = (_round_half_away(x1), _round_half_away(y1), _round_half_away(x2), _round_half_away(y2))
            bh = f32(max(eh - sh + 1, 1)) / f32(ph)
            bw = f32(max(ew - sw + 1, 1)) / f32(pw)
            i = np.arange(ph, dtype=np.float32)
            j = np.arange(pw, dtype=np.float32)
            hs = np.clip(np.floor(i * bh).astype(np.int64) + sh, 0, h)
            he = np.clip(np.ceil((i + f32(1)) * bh).astype(np.int64) + sh, 0, h)
            ws = np.clip(np.floor(j * bw).astype(np.int64) + sw, 0, w)
            we = np.clip(np.ceil((j + f32(1)) * bw).astype(np.int64) + sw, 0, w)
            mh, mw = int(max((he - hs).max(), 0)), int(max((we - ws).max(), 0))
            out = torch.full((c, ph, pw), float('-inf'), dtype=feat.dtype)
            for dy in range(mh):
                rows = torch.from_numpy(np.minimum(hs + dy, h - 1))
                rok = torch.from_numpy(hs + dy < he)
                for dx in range(mw):
                    cols = torch.from_numpy(np.minimum(ws + dx, w - 1))
                    cok = torch.from_numpy(ws + dx < we)
                    v = feat[b][:, rows][:, :, cols]
                    ok = (rok[:, None] & cok[None, :])[None]
                    out = torch.where(ok, torch.maximum(out, v), out)
            outs.append(torch.where(torch.isinf(out), torch.zeros_like(out), out))
    return torch.stack(outs, dim=0)


def resnet_encoder(p, x, n_filters, use_bn, training, pre):
    """src/networks.py:232-268 (ResNetEncoder.forward)."""
    layers = [conv_block(p, pre + 'conv1', x, 2, 'leaky_relu', use_bn, training)]
    y = F.max_pool2d(layers[-1], 3, 2, 1)
    for level in range(2, len(n_filters) + 1):
        y = resnet_stage(p, pre + 'blocks%d' % level, y, 1 if level == 2 else 2, 2, use_bn, training)
        layers.append(y)
    return layers[-1], layers[:-1]


def mlp_encoder(p, points, pre, n_layers=6):
    """src/networks.py:1007-1067: 6 x (Linear + LeakyReLU(0.2)) incl. the last."""
    x = points
    for i in range(n_layers):
        x = F.leaky_relu(F.linear(x, p[pre + 'mlp.%d.fully_connected.weight' % i],
                                  p[pre + 'mlp.%d.fully_connected.bias' % i]), LEAKY_SLOPE)
    return x


def radarnet_forward(p, image, points, boxes_list, patch_size, n_filters_image=(32, 64, 128, 128, 128),
                     n_neuron_latent=128, training=False, return_logits=True, roi_pool=roi_pool):
    """src/radarnet_model.py:102-124 + src/networks.py:1203-1256."""
    ph, pw = patch_size
    lat_h, lat_w = int(ph // 32.0), int(pw // 32.0)
    scales = [1 / 2.0, 1 / 4.0, 1 / 8.0, 1 / 16.0, 1 / 32.0, 1 / 64.0, 1 / 128.0]
    latent_img, skips_img = resnet_encoder(p, image, n_filters_image, True, training,
                                           'encoder.encoder_image.')
    latent_pooled = roi_pool(latent_img, boxes_list, 1 / 32.0, (lat_h, lat_w))
    skips = [roi_pool(s, boxes_list, scales[i], (int(ph * scales[i]), int(pw * scales[i])))
             for i, s in enumerate(skips_img)]
    lat_d = mlp_encoder(p, points, 'encoder.encoder_depth.')
    lat_d = lat_d.view(points.shape[0], n_neuron_latent, -1, lat_w)
    latent = torch.cat([latent_pooled, lat_d], dim=1)
    logits = multiscale_decoder(p, latent, skips, patch_size, True, training)
    return logits if return_logits else torch.sigmoid(logits)
