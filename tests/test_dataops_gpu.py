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
