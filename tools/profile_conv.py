"""Runs single convolutions of the FusionNet decoder through the C-ABI (for ncu captures)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'radar-camera-fusion-depth_b200'))
from rcfd import ops  # noqa: E402

H, W, B = 352, 704, 8
dev = torch.device('cuda:0')
which = sys.argv[1] if len(sys.argv) > 1 else 'deconv0_up'
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
bf = torch.bfloat16
if which == 'deconv0_up':        # 64 -> 32, fused 2x up-sample, 352x704 output
    x = torch.randn(B, H // 2, W // 2, 64, device=dev).to(bf)
    w32 = torch.randn(32, 64, 3, 3, device=dev) * 0.05
    w = ops.pack_weight(w32, bf)
    wup = ops.pack_upconv2x_weight(w32, bf) if os.environ.get('RCFD_NO_UP2X') is None else None
    eng = int(os.environ.get('RCFD_ENGINE', '0'))
    run = lambda: ops.conv2d(x, w, 32, 3, 1, in_size=(H, W), weight_up2x=wup, engine=eng)
elif which == 'deconv0_conv':    # 32 -> 32 at 352x704
    x = torch.randn(B, H, W, 32, device=dev).to(bf)
    w = ops.pack_weight(torch.randn(32, 32, 3, 3, device=dev) * 0.05, bf)
    eng = int(os.environ.get('RCFD_ENGINE', '0'))
    run = lambda: ops.conv2d(x, w, 32, 3, 1, engine=eng)
elif which == 'deconv1_up_lowres':   # 64 -> 64 at 176x352 (plain 3x3 at that resolution)
    x = torch.randn(B, H // 2, W // 2, 64, device=dev).to(bf)
    w = ops.pack_weight(torch.randn(64, 64, 3, 3, device=dev) * 0.05, bf)
    run = lambda: ops.conv2d(x, w, 64, 3, 1)
elif which == 'blocks4':         # 256 -> 256 at 22x44
    x = torch.randn(B, 22, 44, 256, device=dev).to(bf)
    w = ops.pack_weight(torch.randn(256, 256, 3, 3, device=dev) * 0.02, bf)
    run = lambda: ops.conv2d(x, w, 256, 3, 1)
elif which == 'wgrad_deconv0_up':
    x = torch.randn(B, H // 2, W // 2, 64, device=dev).to(bf)
    dy = torch.randn(B, H, W, 32, device=dev).to(bf)
    eng = int(os.environ.get('RCFD_ENGINE', '0'))
    run = lambda: ops.conv2d_wgrad(x, dy, 3, 1, in_size=(H, W), engine=eng)
elif which == 'wgrad_deconv0_conv':      # 32 -> 32 at 352x704
    x = torch.randn(B, H, W, 32, device=dev).to(bf)
    dy = torch.randn(B, H, W, 32, device=dev).to(bf)
    eng = int(os.environ.get('RCFD_ENGINE', '0'))
    run = lambda: ops.conv2d_wgrad(x, dy, 3, 1, engine=eng)
elif which == 'wgrad_output0':           # 32 -> 1 (d(logit) stored with 16 channels) at 352x704
    x = torch.randn(B, H, W, 32, device=dev).to(bf)
    dy = torch.randn(B, H, W, 16, device=dev).to(bf)
    eng = int(os.environ.get('RCFD_ENGINE', '0'))
    run = lambda: ops.conv2d_wgrad(x, dy, 3, 1, engine=eng)
elif which == 'wgrad_deconv1_conv':      # (64 | 32) -> 64 at 176x352
    x = torch.randn(B, H // 2, W // 2, 64, device=dev).to(bf)
    x1 = torch.randn(B, H // 2, W // 2, 32, device=dev).to(bf)
    dy = torch.randn(B, H // 2, W // 2, 64, device=dev).to(bf)
    eng = int(os.environ.get('RCFD_ENGINE', '0'))
    run = lambda: ops.conv2d_wgrad(x, dy, 3, 1, x1=x1, engine=eng)
elif which == 'wgrad_deconv1_up':        # 64 -> 64 behind a 2x up-sampling, 176x352 output
    x = torch.randn(B, H // 4, W // 4, 64, device=dev).to(bf)
    dy = torch.randn(B, H // 2, W // 2, 64, device=dev).to(bf)
    eng = int(os.environ.get('RCFD_ENGINE', '0'))
    run = lambda: ops.conv2d_wgrad(x, dy, 3, 1, in_size=(H // 2, W // 2), engine=eng)
elif which == 'dgrad_output0':           # d(logit) (16 channels) -> 32 at 352x704
    dy = torch.randn(B, H, W, 16, device=dev).to(bf)
    wd = ops.pack_weight(torch.randn(1, 32, 3, 3, device=dev) * 0.05, bf, dgrad=True, pad_to=16)
    run = lambda: ops.conv2d(dy, wd, 32, 3, 1, pad=1, out_size=(H, W))
elif which == 'maxpool_bwd':
    x = torch.randn(B, H // 2, W // 2, 32, device=dev).to(bf)
    d = torch.randn(B, H // 4, W // 4, 32, device=dev).to(bf)
    run = lambda: ops.maxpool3x3s2_bwd(x, d)
elif which == 'upsample_bwd':
    d = torch.randn(B, H, W, 64, device=dev).to(bf)
    run = lambda: ops.upsample_nearest_bwd(d, (H // 2, W // 2))
else:
    raise SystemExit('unknown case')
for _ in range(reps):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    run()
e1.record()
torch.cuda.synchronize()
print(which, 'ms/launch', e0.elapsed_time(e1) / reps)
