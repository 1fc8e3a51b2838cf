"""GPU parity of the data-path kernels against the oracle and the fixtures written by the
unmodified reference.  Through the C-ABI (rcfd.ops -> librcfd_b200.so)."""
import numpy as np
import pytest
import torch

from helpers import load_golden

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda:0')


def test_batched_transforms_kernel_matches_reference_fixture():
    """fusionnet_transforms.Transforms on CUDA tensors (rcfd_transform_batch: one fused apply kernel for the batch) vs
    tests/golden/transforms_5x18x26.npz, written by the reference's per-sample loop on the same seeded inputs; the random
    draws are made on the CPU generator so that they are the reference's.  Range maps (flips): bit exact.  Image: bit
    exact except where the contrast partner -- the mean of the grey image, summed exactly (float64) by the kernel and
    in float32 by torch.mean -- moves a blended value across an integer: at most one grey level on isolated pixels."""
    import fusionnet_transforms
    g = load_golden('transforms_5x18x26')
    seed = int(g['meta'][0])
    cfg = dict(random_brightness=[0.8, 1.2], random_contrast=[0.8, 1.2], random_saturation=[0.8, 1.2],
               random_flip_type=['horizontal', 'vertical'])
    for tag, scale, rng in (('u8_01', 255.0, [0, 1]), ('f_pm1', 1.0, [-1, 1])):
        gen = torch.Generator().manual_seed(seed)
        img = torch.rand(5, 3, 18, 26, generator=gen) * scale
        if scale > 1.0:
            img = img.round()
        maps = [torch.rand(5, 1, 18, 26, generator=gen) * 50, torch.rand(5, 2, 18, 26, generator=gen)]
        t = fusionnet_transforms.Transforms(normalized_image_range=rng, rand_device='cpu', **cfg)
        torch.manual_seed(seed + 1)
        (oi,), om = t.transform([img.to(DEV)], [m.to(DEV) for m in maps], random_transform_probability=0.9)
        assert np.array_equal(om[0].cpu().numpy(), g[tag + '_map0']) and np.array_equal(om[1].cpu().numpy(), g[tag + '_map1'])
        ref = g[tag + '_image']
        diff = np.abs(oi.cpu().numpy() - ref)
        level = (1.0 / 255.0 if rng == [0, 1] else 2.0 / 255.0) if scale > 1.0 else 1e-6
        print(tag, 'max |diff|', diff.max(), 'pixels differing', int((diff > 0).sum()), 'of', diff.size)
        assert diff.max() <= level * 1.0001
        assert (diff > 1e-6).mean() < 0.01
    # no augmentation at all: pure normalisation, every variant, no range maps -> a bare list like the reference
    x = (torch.rand(2, 3, 20, 36) * 255).round()
    for rng, fn in (([0, 255], lambda v: v), ([0, 1], lambda v: v / 255.0), ([-1, 1], lambda v: 2.0 * (v / 255.0) - 1.0)):
        out = fusionnet_transforms.Transforms(normalized_image_range=rng).transform([x.to(DEV)], random_transform_probability=0.0)
        assert isinstance(out, list) and torch.equal(out[0].cpu(), fn(x))
    # six range maps (two kernel calls), flips only, on-device draws: every map is flipped consistently with the image
    t = fusionnet_transforms.Transforms(normalized_image_range=[0, 255], random_flip_type=['horizontal', 'vertical'])
    maps = [torch.rand(4, 1, 20, 36, device=DEV) for _ in range(6)]
    (oi,), om = t.transform([(torch.rand(4, 3, 20, 36) * 200 + 20).to(DEV)], maps, random_transform_probability=1.0)
    for m_in, m_out in zip(maps, om):
        for b in range(4):
            cands = [m_in[b], m_in[b].flip(-1), m_in[b].flip(-2), m_in[b].flip(-1).flip(-2)]
            assert any(torch.equal(m_out[b], c) for c in cands)
    which = [[torch.equal(om[0][b], c) for c in (maps[0][b], maps[0][b].flip(-1), maps[0][b].flip(-2), maps[0][b].flip(-1).flip(-2))].index(True)
             for b in range(4)]
    for m_in, m_out in zip(maps[1:], om[1:]):
        for b in range(4):
            c = (m_in[b], m_in[b].flip(-1), m_in[b].flip(-2), m_in[b].flip(-1).flip(-2))[which[b]]
            assert torch.equal(m_out[b], c)


def test_decode_crop_and_encode_kernels():
    """rcfd_decode_crop / rcfd_encode_u16 vs the reference's value codec restated in numpy (load_image / load_depth /
    load_response: raster / multiplier, <= 0 -> 0; save_*: uint32(v * multiplier) as 16 bits; src/data_utils.py:167-335)
    and its crop (src/datasets.py:101-109), bit exact; depth and response land in the two channels of input_depth."""
    from rcfd import data, ops
    rng = np.random.RandomState(5)
    n, h0, w0, oh, ow = 3, 37, 61, 24, 40
    img = rng.randint(0, 256, (n, h0, w0, 3)).astype(np.uint8)
    z = (rng.rand(n, h0, w0) * 65535 * (rng.rand(n, h0, w0) < 0.5)).astype(np.uint16)
    r = (rng.rand(n, h0, w0) * 16384).astype(np.uint16)
    origin = np.stack([rng.randint(0, h0 - oh + 1, n), rng.randint(0, w0 - ow + 1, n)], 1).astype(np.int32)
    d_img, d_z, d_r = torch.from_numpy(img).to(DEV), torch.from_numpy(z.view(np.int16)).to(DEV), torch.from_numpy(r.view(np.int16)).to(DEV)
    d_o = torch.from_numpy(origin).to(DEV)
    crop = lambda a, i: a[i, origin[i, 0]:origin[i, 0] + oh, origin[i, 1]:origin[i, 1] + ow]
    out = ops.decode_crop(d_img, 1.0, (oh, ow), d_o).cpu().numpy()
    for i in range(n):
        assert np.array_equal(out[i], np.transpose(crop(img, i).astype(np.float32), (2, 0, 1)))
    both = torch.full((n, 2, oh, ow), -1.0, device=DEV)
    ops.decode_crop(d_z, 256.0, (oh, ow), d_o, out=both, out_channel=0)
    ops.decode_crop(d_r, 2.0 ** 14, (oh, ow), d_o, out=both, out_channel=1)
    both = both.cpu().numpy()
    for i in range(n):
        ref_z = crop(z, i).astype(np.float32) / 256.0
        ref_z[ref_z <= 0] = 0.0
        assert np.array_equal(both[i, 0], ref_z) and np.array_equal(both[i, 1], crop(r, i).astype(np.float32) / 2 ** 14)
    # no crop: whole rasters
    assert np.array_equal(ops.decode_crop(d_z, 256.0).cpu().numpy()[:, 0], z.astype(np.float32) / 256.0)
    # encode: save_depth / save_response quantisation, and the round trip through file precision
    v = torch.rand(2, 1, 9, 13) * 300.0          # beyond 255.99: the 16-bit store wraps like the PNG does
    enc = ops.encode_u16(v.to(DEV), 256.0).cpu().numpy()
    assert np.array_equal(enc, (np.uint32(v.numpy() * np.float32(256.0)) & 0xffff).astype(np.uint16))
    # the five tensors of a training batch from on-disk sample types == the float batch in file precision
    image = (torch.rand(2, 3, 16, 24) * 255).round()
    maps = [torch.rand(2, 1, 16, 24) * 80 * (torch.rand(2, 1, 16, 24) < 0.5) for _ in range(4)]
    raw = data.encode_raw_batch(image, *maps)
    dec = data.decode_fusionnet_batch([t.pin_memory() for t in raw], DEV)
    assert torch.equal(dec[0].cpu(), image)
    for got, want in zip(dec[1:], maps):
        assert torch.equal(got.cpu(), torch.floor(want * 256.0) / 256.0)
