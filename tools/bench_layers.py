"""Times single layers of the FusionNet step through the C-ABI (CUDA events, L2 flushed between launches).

    python tools/bench_layers.py [case ...]          (no arguments: every case)
    RCFD_OPT=key=value,key=value python tools/bench_layers.py ...   (rcfd_set_option knobs)

A case is  name = (kind, cin, cout, h, w, extras): conv forward (optionally with the BatchNorm statistics epilogue,
a second concat source or the fused 2x up-sampling), wgrad, BN backward passes and the layout kernels, at batch 8.
Prints microseconds per launch plus the achieved TFLOP/s and GB/s of the algorithmic work."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'radar-camera-fusion-depth_b200'))
from rcfd import ops  # noqa: E402

B = int(os.environ.get('RCFD_BATCH', '8'))
dev = torch.device('cuda:0')
bf = torch.bfloat16

CASES = {
    # name: (kind, cin, cout, h_out, w_out, dict)
    'stem_image': ('conv', 16, 32, 176, 352, dict(k=4, pad=2, stats=True)),
    'stem_depth': ('conv', 16, 16, 176, 352, dict(k=4, pad=2, stats=True)),
    'b2_img': ('conv', 64, 64, 88, 176, dict(stats=True)),
    'b2_img_nostats': ('conv', 64, 64, 88, 176, dict()),
    'b2_dep': ('conv', 32, 32, 88, 176, dict(stats=True)),
    'b3_img': ('conv', 128, 128, 44, 88, dict(stats=True)),
    'b3_img_s2': ('conv', 64, 128, 44, 88, dict(stats=True, stride=2)),
    'b4_img': ('conv', 256, 256, 22, 44, dict(stats=True)),
    'b5_img': ('conv', 256, 256, 11, 22, dict(stats=True)),
    'b6_img': ('conv', 256, 256, 6, 11, dict(stats=True)),
    'b6_dep': ('conv', 128, 128, 6, 11, dict(stats=True)),
    'b6_img_nostats': ('conv', 256, 256, 6, 11, dict()),
    'fuse4': ('conv', 128, 512, 22, 44, dict(k=1, stats=True)),
    'dec4_conv': ('conv', 256, 256, 22, 44, dict(stats=True, c1=256)),
    'dec3_conv': ('conv', 128, 128, 44, 88, dict(stats=True, c1=128)),
    'dec2_conv': ('conv', 64, 64, 88, 176, dict(stats=True, c1=64)),
    'dec1_up': ('conv', 64, 64, 176, 352, dict(stats=True, up=True)),
    'dec1_conv': ('conv', 64, 64, 176, 352, dict(stats=True, c1=32)),
    'dec0_up': ('conv', 64, 32, 352, 704, dict(stats=True, up=True)),
    'dec0_up_nostats': ('conv', 64, 32, 352, 704, dict(up=True)),
    'dec0_conv': ('conv', 32, 32, 352, 704, dict(stats=True)),
    'dec0_conv_nostats': ('conv', 32, 32, 352, 704, dict()),
    'out0': ('conv', 32, 1, 352, 704, dict(head=True)),
    'wg_dec0_conv': ('wgrad', 32, 32, 352, 704, dict()),
    'wg_dec0_up': ('wgrad', 64, 32, 352, 704, dict(up=True)),
    'wg_dec1_conv': ('wgrad', 64, 64, 176, 352, dict(c1=32)),
    'wg_b2_img': ('wgrad', 64, 64, 88, 176, dict()),
    'wg_b3_img': ('wgrad', 128, 128, 44, 88, dict()),
    'wg_b4_img': ('wgrad', 256, 256, 22, 44, dict()),
    'wg_b6_img': ('wgrad', 256, 256, 6, 11, dict()),
    'wg_stem': ('wgrad', 16, 32, 176, 352, dict(k=4, pad=2)),
    'wg_stem_depth': ('wgrad', 16, 16, 176, 352, dict(k=4, pad=2)),
    'bnbwd_dec0': ('bnbwd', 32, 32, 352, 704, dict()),
    'bnbwd_b2': ('bnbwd', 64, 64, 88, 176, dict()),
    'bnbwd_b4': ('bnbwd', 256, 256, 22, 44, dict()),
    'bnbwd_b6': ('bnbwd', 256, 256, 6, 11, dict()),
    'bnfwd_dec0': ('bnfwd', 32, 32, 352, 704, dict()),
    's2d': ('s2d', 3, 16, 352, 704, dict()),
    'head_bwd': ('headbwd', 1, 16, 352, 704, dict()),
    'outlier': ('outlier', 1, 1, 352, 704, dict()),
    'maxpool_bwd': ('poolbwd', 32, 32, 176, 352, dict()),
    'maxpool_bwd_idx': ('poolbwd', 32, 32, 176, 352, dict(idx=True)),
    'maxpool_fwd_idx': ('poolfwd', 32, 32, 176, 352, dict(idx=True)),
    'maxpool_fwd': ('poolfwd', 32, 32, 176, 352, dict()),
}


def build(kind, cin, cout, h, w, o):
    k = o.get('k', 3)
    stride = o.get('stride', 1)
    if kind == 'conv':
        hin, win = h * stride, w * stride
        c1 = o.get('c1', 0)
        if o.get('up'):
            x = torch.randn(B, h // 2, w // 2, cin, device=dev).to(bf)
        else:
            x = torch.randn(B, hin, win, cin, device=dev).to(bf)
        x1 = torch.randn(B, hin, win, c1, device=dev).to(bf) if c1 else None
        w32 = torch.randn(cout, cin + c1, k, k, device=dev) * 0.05
        wt = ops.pack_weight(w32, bf)
        wup = ops.pack_upconv2x_weight(w32, bf) if o.get('up') else None
        stats = (torch.zeros(cout, device=dev, dtype=torch.float64), torch.zeros(cout, device=dev, dtype=torch.float64)) \
            if o.get('stats') else None
        kw = dict(x1=x1, stats=stats, weight_up2x=wup, pad=o.get('pad'), engine=int(os.environ.get('RCFD_ENGINE', '0')))
        if o.get('up'):
            kw['in_size'] = (h, w)
        if o.get('k') == 4:
            kw['out_size'] = (h, w)
        if o.get('head'):
            kw.update(act=ops.ACT_DEPTH_HEAD, act_params=(1.0, 0.01), out_f32=True)
        out = [None]

        def run():
            out[0] = ops.conv2d(x, wt, cout, k, stride, out=out[0], **kw)
        flops = 2.0 * B * h * w * cout * k * k * (cin + c1)
        byts = (x.numel() + (x1.numel() if c1 else 0) + B * h * w * cout) * 2
        return run, flops, byts
    if kind == 'wgrad':
        c1 = o.get('c1', 0)
        x = torch.randn(B, h // 2 if o.get('up') else h, w // 2 if o.get('up') else w, cin, device=dev).to(bf)
        x1 = torch.randn(B, h, w, c1, device=dev).to(bf) if c1 else None
        dy = torch.randn(B, h, w, cout, device=dev).to(bf)
        kw = dict(x1=x1, pad=o.get('pad'))
        if o.get('up'):
            kw['in_size'] = (h, w)
        run = lambda: ops.conv2d_wgrad(x, dy, k, 1, **kw)
        return run, 2.0 * B * h * w * cout * k * k * (cin + c1), (x.numel() + (x1.numel() if c1 else 0) + dy.numel()) * 2
    if kind in ('bnbwd', 'bnfwd'):
        y = torch.randn(B, h, w, cout, device=dev).to(bf)
        dz = torch.randn(B, h, w, cout, device=dev).to(bf)
        sc = torch.rand(cout, device=dev) + 0.5
        sh = torch.randn(cout, device=dev)
        mu = torch.randn(cout, device=dev)
        inv = torch.rand(cout, device=dev) + 0.5
        dg, db = torch.empty(cout, device=dev), torch.empty(cout, device=dev)
        if kind == 'bnfwd':
            return (lambda: ops.bn_act(y, sc, sh, ops.ACT_LEAKY)), 0.0, 2 * y.numel() * 2
        return (lambda: ops.bn_act_bwd(dz, y, sc, sh, mu, inv, ops.ACT_LEAKY, dg, db)), 0.0, 5 * y.numel() * 2
    if kind == 's2d':
        x = torch.rand(B, 3, h, w, device=dev)
        return (lambda: ops.nchw_to_s2d(x, bf, 16)), 0.0, x.numel() * 4 + B * (h // 2) * (w // 2) * 16 * 2
    if kind == 'headbwd':
        d = torch.rand(B, h, w, 1, device=dev) + 1.0
        dd = torch.randn(B, h, w, 1, device=dev)
        return (lambda: ops.depth_head_bwd(dd, d, 1.0, 0.01, bf, cpad=16)), 0.0, B * h * w * (8 + 32)
    if kind == 'outlier':
        d = torch.rand(B, 1, h, w, device=dev) * (torch.rand(B, 1, h, w, device=dev) < 0.3)
        return (lambda: ops.outlier_removal(d, 7, 1.5)), 0.0, B * h * w * 12
    if kind in ('poolbwd', 'poolfwd'):
        x = torch.randn(B, h, w, cin, device=dev).to(bf)
        d = torch.randn(B, h // 2, w // 2, cin, device=dev).to(bf)
        if kind == 'poolfwd':
            return (lambda: ops.maxpool3x3s2_idx(x)) if o.get('idx') else (lambda: ops.maxpool3x3s2(x)), 0.0, (x.numel() + d.numel()) * 2
        if o.get('idx'):
            _, idx = ops.maxpool3x3s2_idx(x)
            return (lambda: ops.maxpool3x3s2_bwd_idx(d, idx, (h, w))), 0.0, (x.numel() + d.numel()) * 2 + d.numel()
        return (lambda: ops.maxpool3x3s2_bwd(x, d)), 0.0, (2 * x.numel() + d.numel()) * 2
    raise SystemExit('unknown kind ' + kind)


if __name__ == '__main__':
    names = sys.argv[1:] or list(CASES)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    reps = int(os.environ.get('RCFD_REPS', '5'))
    for name in names:
        run, flops, byts = build(*CASES[name])
        try:
            for _ in range(2):
                run()
            tot = 0.0
            for _ in range(reps):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                run()
                b.record()
                torch.cuda.synchronize()
                tot += a.elapsed_time(b)
            us = tot / reps * 1e3
            print('%-20s %9.1f us  %7.1f TF/s  %7.1f GB/s' % (name, us, flops / us / 1e6, byts / us / 1e3), flush=True)
        except Exception as e:  # noqa: BLE001
            print('%-20s FAILED %s' % (name, e), flush=True)
