#!/usr/bin/env python
"""
bench.py -- FusionNet depth-maps/s @352x704 on N B200s (BASELINE.json metric) and the other BASELINE configs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--mode train|infer|radarnet] [--batch B]
                    [--precision bf16|fp32|bf16x3|bf16x6] [--impl ours|reference|cudnn]

One "step" is one pass of the hot path over one synthetic batch:
  train    (default, BASELINE configs[1]): forward + masked-L1 loss + backward + Adam, batch 8 / GPU, bf16;
  infer    (configs[4])                  : eval-mode forward, batch B / GPU;
  radarnet (configs[2])                  : RadarNet stage-1 forward + S2 scatter, 16 images x 64 radar points / GPU.
`value` is whole-job throughput with inputs resident in HBM; `e2e` is the same step through the public API with
pinned-host inputs copied H2D and the result (loss / depth maps / depth + response maps) read back D2H inside the
timed region.  The FusionNet e2e arms are measured with both input forms of the public API -- float32 tensors
(train_step_graphed / forward_graphed) and the reference's on-disk sample types, uint8 RGB + uint16 maps decoded on
the device (train_step_graphed_raw / forward_graphed_raw, 2.5x fewer PCIe bytes) -- `e2e` is the float32 form on one GPU
and the on-disk form when several ranks share the host (measured: 10 468 vs 10 303 maps/s on 8 GPUs); the other one is
reported under `e2e.other_input_form`.  N > 1: launched by torchrun, one rank per GPU, gradients all-reduced over NCCL (train) or independent
replicas (infer, radarnet); weak scaling.

The default run prints ONE JSON line for the train step; at N = 1 that line also carries
  `roofline`     the time-dominant kernel of the step (found by timing every C-ABI call of one step alone with CUDA
                 events, rcfd/census.py), algorithmic and executed FLOP/s against the measured peak, plus `kernels`
                 (top kernels by share) and `step_breakdown`;
  `cpu_baseline` the reference's CPU path (oracle port) timed on the box's host cores at the SAME batch;
  `gpu_baseline` PyTorch eager + cuDNN running the reference graph on this GPU (fp32/TF32 and autocast bf16 +
                 channels_last; tools/eager_fusionnet.py) -- the library path the reference itself would use here;
  `also`         the same measurement object for the other BASELINE configs (infer batch 8, radarnet 16 x 64, and the inference sweep at batch 1 / 32 / 128);
  `parity`       the recorded deviation of this arithmetic from the fp32 oracle (profiles/r2_bf16_deviation.json).

--impl reference times the reference's own CPU path (the oracle port: same algorithm, fp32, torch CPU ops on the host
threads a probe finds fastest) on the same config; --impl cudnn prints the eager + cuDNN arm alone.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, 'radar-camera-fusion-depth_b200'))

import torch  # noqa: E402

H, W = 352, 704
K_POINTS = 64
PATCH = (352, 288)
FWD_GFLOP = 57.10            # per depth map, SURVEY.md 8d / BASELINE.md section 2
TRAIN_GFLOP = 170.5
ACT_MB_BF16 = 240.8          # conv activation traffic per depth map, bf16 (fwd)
RADAR_GFLOP = 14.35 + 8.20 * K_POINTS          # per image at K = 64

METRIC = {'train': ('FusionNet depth-maps/sec @352x704', 'depth-maps/s'),
          'infer': ('FusionNet depth-maps/sec @352x704', 'depth-maps/s'),
          'radarnet': ('RadarNet stage-1 images/sec @352x704, 64 radar points per image', 'images/s')}


def workload_config(mode, batch, world):
    """The `config` object: identical for every arm (--impl ours | reference | cudnn) of the same workload."""
    name = {'train': 'FusionNet training step (fwd + masked-L1 + bwd + Adam), 352x704, batch %d per GPU (BASELINE configs[1])',
            'infer': 'FusionNet eval forward, 352x704, batch %d per GPU (BASELINE configs[4])',
            'radarnet': 'RadarNet stage-1 forward + S2 scatter, 352x704, %d images x 64 radar points per GPU, patch 352x288 '
                        '(BASELINE configs[2])'}[mode] % batch
    return {'workload': name, 'batch_per_gpu': batch, 'global_batch': world * batch, 'parallelism': 'dp%d' % world,
            'l2': 'no flush needed: the per-step working set (> 1 GB of activations) exceeds the 126 MB L2'}


def read_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d['hbm_gbs'], tc_burst=d['bf16_tflops'], tc_sustained=d['bf16_tflops_sustained'],
                    source='measured (MEASURED_PEAKS.json)')
    return dict(hbm=6650.0, tc_burst=1590.0, tc_sustained=1400.0, source='fallback (B200_PROFILING.md)')


def read_parity(precision):
    p = os.path.join(ROOT, 'profiles', 'r2_bf16_deviation.json')
    if not os.path.exists(p):
        return None
    d = json.load(open(p))
    step = d.get('train_step_352x704_b2', {})
    out = {'source': 'profiles/r2_bf16_deviation.json (tests/test_tc_parity_gpu.py, vs the fp32 CPU oracle)',
           'tolerance': 'north star 1e-3 relative: met by the tensor-core parity modes bf16x3 / bf16x6 and by fp32; '
                        'the bf16 fast mode is NOT inside it (deviation below)'}
    key = precision if precision in step else None
    if key:
        out['train_step_352x704_b2'] = step[key]
    if 'bf16x6' in step and precision != 'bf16x6':
        out['train_step_352x704_b2_parity_mode_bf16x6'] = step['bf16x6']
    for k in ('config1_bf16x3', 'config1_bf16x6', 'radarnet_352x704_k64'):
        if k in d:
            out[k] = d[k]
    return out


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '50'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace('.', '').isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i] == 'Active' for r in self.rows)]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(sm)}


def synthetic_batch(batch, seed):
    """SURVEY 8d inputs (cheap generator: uniform image, banded quasi-dense radar depth)."""
    g = torch.Generator().manual_seed(1000 + seed)
    image = torch.rand(batch, 3, H, W, generator=g)
    depth = torch.zeros(batch, 2, H, W)
    for b in range(batch):
        for _ in range(64):
            x = int(torch.randint(16, W - 16, (1,), generator=g))
            y = int(torch.randint(48, H - 16, (1,), generator=g))
            z = float(torch.rand(1, generator=g)) * 79 + 1
            depth[b, 0, y - 48:y + 16, x - 16:x + 16] = z
    depth[:, 1] = torch.where(depth[:, 0] > 0, torch.rand(batch, H, W, generator=g) * 0.5 + 0.5, torch.zeros(()))
    gt = (torch.rand(batch, 1, H, W, generator=g) * 79 + 1) * (torch.rand(batch, 1, H, W, generator=g) < 0.30)
    lidar = (torch.rand(batch, 1, H, W, generator=g) * 79 + 1) * (torch.rand(batch, 1, H, W, generator=g) < 0.02)
    return image, depth, gt.float(), lidar.float()


def radarnet_batch(n_img, seed):
    """configs[2] inputs: images 3 x 352 x 704, 64 radar points each (x, y, z) and their column boxes
    (reference src/radarnet_main.py:980-990: x +- patch_width / 2 in edge-padded pixel coordinates)."""
    from rcfd import synth
    pad = PATCH[1] // 2
    g = torch.Generator().manual_seed(2000 + seed)
    images = torch.rand(n_img, 3, H, W, generator=g)
    pts, boxes = [], []
    for b in range(n_img):
        pt = synth.radar_points(K_POINTS, H, W, seed * 1000 + b)
        pt[:, 0] += pad
        pts.append(pt)
        boxes.append(torch.stack([pt[:, 0] - pad, torch.zeros(K_POINTS), pt[:, 0] + pad,
                                  torch.full((K_POINTS,), float(H))], 1))
    return images, torch.stack(pts), torch.stack(boxes)


# ----------------------------------------------------------------------------- reference arm (CPU oracle port)
def _oracle_state(seed=0):
    from rcfd import synth
    import networks  # parameter containers only (CPU)
    cfg = synth.CANONICAL_FUSIONNET
    enc = networks.FusionNetEncoder(18, 3, 2, cfg['n_filters_encoder_image'], cfg['n_filters_encoder_depth'],
                                    'kaiming_uniform', 'leaky_relu', True, 'weight_and_project')
    dec = networks.MultiScaleDecoder(256, 1, 1, cfg['n_filters_decoder'], cfg['n_filters_encoder_image'][:-1][::-1] + [0],
                                     'kaiming_uniform', 'leaky_relu', 'linear', True, 'up')
    p = {}
    for k, v in enc.state_dict().items():
        p['encoder.' + k] = v.detach().clone()
    for k, v in dec.state_dict().items():
        p['decoder.' + k] = v.detach().clone()
    synth.fill_state_dict_(p, seed)
    return p


def oracle_step_fn(mode, batch, seed=0):
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    if mode == 'radarnet':
        return oracle_radarnet_step_fn(batch, seed)
    import fusionnet_oracle as fo
    p = _oracle_state(seed)
    image, depth, gt, lidar = synthetic_batch(batch, seed)
    if mode == 'infer':
        def step():
            with torch.no_grad():
                d, _ = fo.fusionnet_forward(p, image, depth)
            return float(d.sum())
        return step
    names = [k for k, v in p.items() if v.is_floating_point() and 'running' not in k]
    for k in names:
        p[k].requires_grad_(True)
    m = [torch.zeros_like(p[k]) for k in names]
    v = [torch.zeros_like(p[k]) for k in names]
    state = {'t': 0}

    def step():
        for k in names:
            p[k].grad = None
        stats = {}
        d, _ = fo.fusionnet_forward(p, image, depth, training=True, new_stats=stats)
        loss = fo.fusionnet_loss(d, fo.outlier_removal(gt, 7, 1.5), lidar, 2.0, 'l1')
        loss.backward()
        state['t'] += 1
        with torch.no_grad():
            fo.adam_step([p[k] for k in names], [p[k].grad for k in names], m, v, state['t'])
            for k, val in stats.items():
                p[k].copy_(val)
        return float(loss.detach())
    return step


def oracle_radarnet_step_fn(n_img, seed=0):
    import radarnet_oracle as ro
    import scatter_oracle as so
    import radarnet_model
    from rcfd import synth
    m = radarnet_model.RadarNetModel(device=torch.device('cpu'), **synth.CANONICAL_RADARNET)
    p = {}
    for k, v in m.encoder.state_dict().items():
        p['encoder.' + k] = v
    for k, v in m.decoder.state_dict().items():
        p['decoder.' + k] = v
    synth.fill_state_dict_(p, seed)
    images, pts, boxes = radarnet_batch(n_img, seed)
    pad = PATCH[1] // 2

    def step():
        tot = 0.0
        with torch.no_grad():
            for b in range(n_img):
                img = torch.nn.functional.pad(images[b:b + 1], (pad, pad, 0, 0), mode='replicate')
                crops = ro.radarnet_forward(p, img, pts[b], [boxes[b]], PATCH, return_logits=False,
                                            roi_pool=ro.roi_pool_vectorised)
                d, r = so.s2_scatter(crops.numpy(), pts[b].numpy(), W, PATCH, compat=True)
                tot += float(r.sum())
        return tot
    return step


def pick_cpu_threads():
    """All host threads the path can USE: time one small forward at a few thread counts and keep
    the fastest (more threads than the convs can feed makes torch's CPU path slower)."""
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    step = oracle_step_fn('infer', 1)
    best, best_t = ncpu, None
    for nt in sorted({ncpu, max(1, ncpu // 2), max(1, ncpu // 4), max(1, ncpu // 8)}, reverse=True):
        torch.set_num_threads(nt)
        step()
        t0 = time.perf_counter()
        step()
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = nt, dt
    torch.set_num_threads(best)
    return best, ncpu


def time_cpu(mode, batch, steps, warmup):
    threads, ncpu = pick_cpu_threads()
    step = oracle_step_fn(mode, batch)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return {'value': batch / dt, 'unit': METRIC[mode][1], 'cores': threads, 'host_cpus': ncpu,
            'os_cpu_count': os.cpu_count(), 'kind': 'port', 'ms_per_step': dt * 1e3,
            'sample': '%d timed %s steps of batch %d at 352x704 (+%d warm-up): the SAME batch as the GPU arm; oracle port '
                      '(functional restatement of the reference, torch CPU fp32), thread count = fastest of a probe over '
                      '{all, 1/2, 1/4, 1/8} of the %d host CPUs' % (steps, mode, batch, warmup, ncpu)}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    batch = args.batch
    if args.mode == 'radarnet':
        batch = min(batch, 1)            # one image x 64 point columns per step: ~0.5 TFLOP of CPU convolution
    steps = args.steps if args.mode != 'radarnet' else min(args.steps, 3)
    warmup = min(args.warmup, 2)
    cpu = time_cpu(args.mode, batch, steps, warmup)
    world = int(os.environ.get('WORLD_SIZE', '1'))
    cfg = workload_config(args.mode, args.batch, world)
    line = {
        'impl': 'reference', 'metric': METRIC[args.mode][0], 'value': cpu['value'], 'unit': METRIC[args.mode][1],
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': cpu['ms_per_step'],
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': cfg, 'timed_steps': steps, 'timed_batch': batch,
        'cpu_baseline': cpu,
        'e2e': {'value': cpu['value'], 'unit': METRIC[args.mode][1], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------- library arm (torch eager + cuDNN on the GPU)
def time_cudnn(mode, batch, dev, steps=5, warmup=3, seed=0):
    """The reference graph in PyTorch eager + cuDNN on this GPU (tools/eager_fusionnet.py): depth-maps/s for
    fp32 (TF32 convolutions allowed = torch's cuDNN default, what the unmodified reference would run) and for
    autocast(bf16) + channels_last."""
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    import eager_fusionnet as ef
    state = {k: v.to(dev) for k, v in _oracle_state(seed).items()}
    tensors = [t.to(dev) for t in synthetic_batch(batch, seed)]
    out = {'kind': 'PyTorch %s eager + cuDNN %s, reference graph restated with the same ATen calls '
                   '(tools/eager_fusionnet.py); none of this repo\'s kernels' % (torch.__version__, torch.backends.cudnn.version()),
           'batch': batch, 'mode': mode, 'unit': 'depth-maps/s', 'steps': steps, 'warmup': warmup,
           'cudnn_allow_tf32': bool(torch.backends.cudnn.allow_tf32)}
    torch.backends.cudnn.benchmark = True
    for variant in ('fp32', 'bf16_cl'):
        step = ef.make_step(state, mode, tensors, variant)
        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out['fp32_tf32_eager' if variant == 'fp32' else 'bf16_autocast_channels_last'] = {'value': batch / ms * 1e3, 'ms_per_step': ms}
        del step
        torch.cuda.empty_cache()
    return out


def run_cudnn(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', '0')))
    torch.cuda.set_device(dev)
    mode = 'train' if args.mode == 'radarnet' else args.mode
    res = time_cudnn(mode, args.batch, dev, steps=args.steps, warmup=max(args.warmup, 3))
    best = max(res['fp32_tf32_eager']['value'], res['bf16_autocast_channels_last']['value'])
    print(json.dumps({'impl': 'cudnn', 'metric': METRIC[mode][0], 'value': best, 'unit': METRIC[mode][1], 'n_gpus': 1,
                      'steps': args.steps, 'warmup': max(args.warmup, 3), 'higher_is_better': True, 'data': 'synthetic',
                      'config': workload_config(mode, args.batch, 1), 'gpu_baseline': res}))


# ----------------------------------------------------------------------------- our arm
class Timer(object):
    def __init__(self, world, dev):
        self.world, self.dev = world, dev

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps, drain=None):
        """K steps bracketed by barrier + synchronize; CUDA events on the current stream; max over ranks."""
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if drain is not None:
            drain()
        e1.record()
        self.barrier()
        ms = e0.elapsed_time(e1)
        if self.world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms / steps


class ResultReader(object):
    """The step's result is copied D2H into one of two pinned buffers and READ by the host while the next step runs
    (the last one inside the timed region too).  Large results (a batch of depth maps) leave the compute stream through
    a device-to-device copy into a staging buffer and cross PCIe on a copy stream, so the next step does not queue
    behind the transfer (what a serving loop does with forward_graphed, whose output buffer the next call overwrites)."""

    def __init__(self, like):
        self.host = [torch.zeros(like.shape, dtype=like.dtype).pin_memory() for _ in range(2)]
        self.ev = [torch.cuda.Event(), torch.cuda.Event()]
        self.i, self.last = 0, None
        self.bytes = like.numel() * like.element_size()
        self.big = self.bytes > (1 << 20)
        if self.big:
            self.dev = [torch.empty_like(like) for _ in range(2)]
            self.copy_stream = torch.cuda.Stream()

    def push(self, t):
        i = self.i
        if self.big:
            self.dev[i % 2].copy_(t, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record()
            self.copy_stream.wait_event(ready)
            with torch.cuda.stream(self.copy_stream):
                self.host[i % 2].copy_(self.dev[i % 2], non_blocking=True)
                self.ev[i % 2].record(self.copy_stream)
        else:
            self.host[i % 2].copy_(t, non_blocking=True)
            self.ev[i % 2].record()
        if i > 0:
            self.drain_one(i - 1)
        self.i = i + 1

    def drain_one(self, j):
        self.ev[j % 2].synchronize()
        self.last = float(self.host[j % 2].flatten()[0])

    def drain(self):
        if self.i > 0:
            self.drain_one(self.i - 1)


def build_workload(args, mode, batch, dev, rank, world):
    """Returns dict(step(inputs), resident, host, result(t) -> tensor copied back, h2d_bytes, model, census_step)."""
    from rcfd import synth, ops
    import net_utils
    torch.manual_seed(0)
    if mode == 'radarnet':
        import radarnet_model
        import radarnet_main
        model = radarnet_model.RadarNetModel(device=dev, **synth.CANONICAL_RADARNET)
        model.set_precision(args.precision)
        model.eval()
        images, pts, boxes = radarnet_batch(batch, rank)
        host = [t.pin_memory() for t in (images, pts, boxes)]
        resident = [t.to(dev) for t in host]

        def step(inputs):
            img, pt, bx = inputs
            if not img.is_cuda:
                img, pt, bx = [t.to(dev, non_blocking=True) for t in (img, pt, bx)]
            with torch.no_grad():
                # the reference's entry point is per image (radarnet_main.forward); forward_batch is the same arithmetic
                # with the encoder run once over the batch and the decoder once over all 16 x 64 point columns
                return radarnet_main.forward_batch(model, img, pt, bx, device=dev)

        def result(out):            # depth (int64, the reference's dtype) + response maps
            return torch.cat([out[0].reshape(-1).view(torch.float32), out[1].reshape(-1)])
        return dict(step=step, resident=resident, host=host, result=result, model=model, opt=None,
                    h2d_bytes=sum(t.numel() * t.element_size() for t in host), eager=step)

    import fusionnet_model
    model = fusionnet_model.FusionNetModel(device=dev, **synth.CANONICAL_FUSIONNET)
    model.set_precision(args.precision)
    model.multistream = args.multistream
    train = mode == 'train'
    opt = outlier = None
    if train:
        model.train()
        if world > 1:
            model.data_parallel()
        from rcfd import optim, parallel
        opt = optim.FusedAdam(model.parameters(), lr=1e-3)      # torch.optim.Adam semantics, one flat kernel
        parallel.use_flat_gradients(model, opt)
        outlier = net_utils.OutlierRemoval(7, 1.5)
    else:
        model.eval()
    host = [t.pin_memory() for t in synthetic_batch(batch, rank)]
    if not train:
        host = host[:2]
    resident = [t.to(dev) for t in host]

    def eager(inputs):
        if train:
            image, depth, gt, lidar = [t if t.is_cuda else t.to(dev, non_blocking=True) for t in inputs]
            out = model.forward(image, depth)
            gt_c = outlier.remove_outliers(gt)
            loss, _ = model.compute_loss(image=image, output_depth=out, ground_truth=gt_c, lidar_map=lidar,
                                         loss_func='l1', w_smoothness=0.0, loss_smoothness_kernel_size=-1,
                                         validity_map_loss_smoothness=None, w_lidar_loss=2.0)
            opt.zero_grad()
            loss.backward()
            opt.step()
            return loss.detach()
        image, depth = [t if t.is_cuda else t.to(dev, non_blocking=True) for t in inputs]
        with torch.no_grad():
            return model.forward(image, depth)

    def step(inputs):
        if not args.graph:
            return eager(inputs)
        if train:       # the graphed entry points stage pinned-host tensors on a copy stream themselves
            return model.train_step_graphed(inputs[0], inputs[1], inputs[2], inputs[3], opt, 2.0, outlier_removal=outlier)
        with torch.no_grad():
            return model.forward_graphed(inputs[0], inputs[1])

    def result(out):
        return out.detach().reshape(-1) if not train else out.detach().reshape(1)
    wl = dict(step=step, resident=resident, host=host, result=result, model=model, opt=opt,
              h2d_bytes=sum(t.numel() * t.element_size() for t in host), eager=eager)
    if not train and args.graph:
        # end-to-end arm of the inference configs: uint8 image + uint16 depth / response from pinned host memory (7 bytes per
        # pixel instead of 20), decoded on the device; the depth map is read back as float32 every batch
        from rcfd import data as rcfd_data
        image, depth = host
        zero = torch.zeros_like(depth[:, 0:1])
        raw = rcfd_data.encode_raw_batch(image * 255.0, depth[:, 0:1], depth[:, 1:2], zero, zero)[:3]
        raw = [t.pin_memory() for t in raw]
        wl['host_raw'] = raw
        wl['h2d_bytes_raw'] = sum(t.numel() * t.element_size() for t in raw)

        def step_raw(inputs):
            with torch.no_grad():
                return model.forward_graphed_raw(inputs)
        wl['step_raw'] = step_raw
    if train and args.graph:
        # end-to-end arm: the batch crosses PCIe in the reference's ON-DISK sample types (uint8 RGB, uint16 16-bit-PNG maps:
        # what rcfd.data.FusionNetRawDataset reads, 11 bytes per pixel instead of 28 as float32) and is decoded by
        # rcfd_decode_crop on the device, the way fusionnet_main.train consumes real data
        from rcfd import data as rcfd_data
        image, depth, gt, lidar = host
        raw = rcfd_data.encode_raw_batch(image * 255.0, depth[:, 0:1], depth[:, 1:2], gt, lidar)[:5]
        raw = [t.pin_memory() for t in raw]
        wl['host_raw'] = raw
        wl['h2d_bytes_raw'] = sum(t.numel() * t.element_size() for t in raw)
        wl['step_raw'] = lambda inputs: model.train_step_graphed_raw(inputs, opt, 2.0, outlier_removal=outlier)
    return wl


def measure(args, mode, batch, dev, rank, world, peaks, want_census, want_cpu):
    from rcfd import _lib, census
    timer = Timer(world, dev)
    wl = build_workload(args, mode, batch, dev, rank, world)
    step, model = wl['step'], wl['model']
    train = mode == 'train'

    # ---- device-resident timing
    for _ in range(args.warmup):
        step(wl['resident'])
    sampler = ClockSampler(dev.index or 0)
    if rank == 0:
        sampler.start()
    ms = timer.timed(lambda: step(wl['resident']), args.steps)
    clocks = sampler.stop() if rank == 0 else None

    # ---- kernels per step: count the C-ABI calls of one eagerly issued step (what the CUDA graph replays)
    l0 = _lib.launch_count
    if mode == 'radarnet' or not args.graph:
        step(wl['resident'])
        launches = _lib.launch_count - l0
    elif train:
        launches = getattr(model, 'last_capture_launches', 0)
    else:
        with torch.no_grad():
            model.forward(wl['resident'][0], wl['resident'][1])
        launches = _lib.launch_count - l0

    # ---- end to end through the public API: pinned host -> device every step, result read back every step
    probe = wl['result'](step(wl['resident']))
    reader = ResultReader(probe)

    def e2e_step():
        reader.push(wl['result'](step(wl['host'])))
    for _ in range(2):
        e2e_step()
    reader.drain()
    ms_e2e = timer.timed(e2e_step, args.steps, drain=reader.drain)
    e2e_inputs, h2d_bytes, e2e_alt = 'float32 tensors (pinned host)', wl['h2d_bytes'], None
    if 'step_raw' in wl:
        # second input form: the on-disk sample types.  Measured: one GPU has PCIe to itself and the float32 form is as fast
        # or faster (the decode kernels sit on the compute stream); with 8 ranks sharing the host the 2.5x smaller copies
        # win (10 468 vs 10 303 maps/s).  The headline e2e uses the form of that regime, the other one is reported beside it.
        def e2e_step_raw():
            reader.push(wl['result'](wl['step_raw'](wl['host_raw'])))
        for _ in range(3):
            e2e_step_raw()
        reader.drain()
        ms_raw = timer.timed(e2e_step_raw, args.steps, drain=reader.drain)
        raw_desc = 'on-disk sample types (uint8 RGB + uint16 maps, pinned host), decoded on the device (rcfd_decode_crop)'
        f32 = {'inputs': e2e_inputs, 'value': world * batch / (ms_e2e * 1e-3), 'ms_per_step': ms_e2e,
               'h2d_bytes_per_step': wl['h2d_bytes']}
        raw = {'inputs': raw_desc, 'value': world * batch / (ms_raw * 1e-3), 'ms_per_step': ms_raw,
               'h2d_bytes_per_step': wl['h2d_bytes_raw']}
        if world > 1:
            ms_e2e, h2d_bytes, e2e_inputs, e2e_alt = ms_raw, wl['h2d_bytes_raw'], raw_desc, f32
        else:
            e2e_alt = raw

    per_unit_gflop = {'train': TRAIN_GFLOP, 'infer': FWD_GFLOP, 'radarnet': RADAR_GFLOP}[mode]
    value = world * batch / (ms * 1e-3)
    line = {
        'metric': METRIC[mode][0], 'value': value, 'unit': METRIC[mode][1],
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': {'bf16': 'bf16', 'fp32': 'f32'}.get(args.precision, args.precision), 'data': 'synthetic',
        'config': workload_config(mode, batch, world),
        'engine': {'precision': args.precision, 'cuda_graph': bool(args.graph) and mode != 'radarnet',
                   'multistream': bool(args.multistream) and mode != 'radarnet'},
        'clocks': clocks,
        'e2e': {'value': world * batch / (ms_e2e * 1e-3), 'unit': METRIC[mode][1], 'ms_per_step': ms_e2e,
                'h2d_bytes_per_step': h2d_bytes, 'd2h_bytes_per_step': reader.bytes, 'inputs': e2e_inputs,
                'other_input_form': e2e_alt},
        'gpu_launches': launches,
        'step_roofline': {'tensor_frac': value / world * per_unit_gflop * 1e9 / (peaks['tc_sustained'] * 1e12),
                          'gflop_per_unit': per_unit_gflop, 'tflops': value / world * per_unit_gflop / 1e3,
                          'peak_tflops_sustained': peaks['tc_sustained'], 'peaks': peaks['source']},
    }
    if mode != 'radarnet':
        line['step_roofline']['hbm_frac_fwd_activations'] = value / world * ACT_MB_BF16 * 1e6 / (peaks['hbm'] * 1e9)

    # ---- roofline of the time-dominant kernel: every C-ABI call of one step timed alone (CUDA events, L2 flushed)
    if want_census and rank == 0:
        ms_flag = getattr(model, 'multistream', False)
        if hasattr(model, 'multistream'):
            model.multistream = False
        with census.record() as rec:
            wl['eager'](wl['resident'])
        torch.cuda.synchronize()
        rows = census.replay(rec.calls, reps=2)
        rec.release()
        if hasattr(model, 'multistream'):
            model.multistream = ms_flag
        kernels = census.by_kernel(rows)
        convs = [k for k in kernels if k['flops'] > 0]
        dom = convs[0]
        total_ms = sum(k['ms'] for k in kernels)
        tf = dom['flops'] / (dom['ms'] * 1e-3) / 1e12
        gbs = dom['bytes'] / (dom['ms'] * 1e-3) / 1e9
        tensor_bound = tf / peaks['tc_burst'] >= gbs / peaks['hbm']
        traffic = None
        tp = os.path.join(ROOT, 'profiles', 'roofline_kernel_traffic.json')
        if os.path.exists(tp):          # dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu --set full)
            tj = json.load(open(tp))
            ent = tj.get('kernels', {}).get('%s|%s|%d' % (dom['kernel'], mode, batch))
            traffic = ent.get('dram_bytes_per_launch') if ent else None
        line['roofline'] = {
            'bound': 'tensor' if tensor_bound else 'hbm',
            'achieved': tf if tensor_bound else gbs, 'peak': peaks['tc_burst'] if tensor_bound else peaks['hbm'],
            'unit': 'TFLOP/s' if tensor_bound else 'GB/s',
            'frac': (tf / peaks['tc_burst']) if tensor_bound else (gbs / peaks['hbm']),
            'traffic': traffic,
            'kernel': dom['kernel'], 'launches_per_step': dom['launches'], 'share_of_step_kernel_time': dom['share'],
            'kernel_ms_per_step': dom['ms'], 'avg_launch_us': dom['ms'] / dom['launches'] * 1e3,
            'algorithmic_gflop_per_step': dom['flops'] / 1e9, 'executed_gflop_per_step': dom['executed_flops'] / 1e9,
            'achieved_tflops': tf, 'executed_tflops': dom['executed_flops'] / (dom['ms'] * 1e-3) / 1e12,
            'algorithmic_mb_per_step': dom['bytes'] / 1e6, 'achieved_gbs': gbs, 'peaks': peaks['source'],
            'how': 'time-dominant conv kernel of the step; every C-ABI call of one eagerly issued step re-issued alone between '
                   'CUDA events on its launch stream, L2 flushed before each (rcfd/census.py); achieved = sum of the algorithmic '
                   'FLOPs of its launches / sum of their durations; peak = burst figure (kernel timed alone)'}
        line['kernels'] = [{'kernel': k['kernel'], 'launches': k['launches'], 'ms': round(k['ms'], 4), 'share': round(k['share'], 4),
                            'tflops': (k['flops'] / (k['ms'] * 1e-3) / 1e12) if k['flops'] else None} for k in kernels[:12]]
        line['step_breakdown'] = {'serialised_kernel_ms': total_ms, 'calls': len(rows), 'by_category': census.by_category(rows)}
    if want_cpu and rank == 0 and world == 1:
        line['cpu_baseline'] = time_cpu(mode, batch if mode != 'radarnet' else 1, 3 if mode != 'radarnet' else 1, 1)
    line['parity'] = read_parity(args.precision)
    # free this workload before the next one
    del wl, step, model
    torch.cuda.empty_cache()
    return line


def run_ours(args):
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    peaks = read_peaks()
    batch = args.batch if args.batch > 0 else {'train': 8, 'infer': 8, 'radarnet': 16}[args.mode]

    if args.profile_step:            # for ncu launch lists: warm up, one more step, nothing else
        wl = build_workload(args, args.mode, batch, dev, rank, world)
        for _ in range(max(args.warmup, 1)):
            wl['step'](wl['resident'])
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_push('profile_step')
        wl['step'](wl['resident'])
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_pop()
        return

    solo = world == 1
    line = measure(args, args.mode, batch, dev, rank, world, peaks, want_census=solo and not args.no_census,
                   want_cpu=not args.no_cpu)
    if solo and args.mode == 'train' and not args.no_also:
        # the other BASELINE configs, same measurement object each (driver-visible in the one JSON line)
        also = {}
        # (BASELINE configs[4]: the inference sweep 1 / 8 / 32 / 128; the roofline census only at batch 8)
        for mode, b in (('infer', 8), ('radarnet', 16), ('infer', 1), ('infer', 32), ('infer', 128)):
            sub = measure(args, mode, b, dev, rank, world, peaks, want_census=not args.no_census and b in (8, 16), want_cpu=False)
            sub.pop('parity', None)
            also['%s_b%d' % (mode, b)] = sub
        line['also'] = also
    if solo and args.mode in ('train', 'infer') and not args.no_gpu_baseline:
        gb = {args.mode: time_cudnn(args.mode, batch, dev)}
        if args.mode == 'train' and not args.no_also:
            gb['infer'] = time_cudnn('infer', 8, dev)
        for k, v in gb.items():
            ours = line['value'] if k == args.mode else line['also']['infer_b8']['value']
            v['ours_over_best_library'] = ours / max(v['fp32_tf32_eager']['value'], v['bf16_autocast_channels_last']['value'])
        line['gpu_baseline'] = gb
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--mode', choices=['train', 'infer', 'radarnet'], default='train')
    ap.add_argument('--batch', type=int, default=0, help='per GPU; default 8 (train, infer) / 16 images (radarnet)')
    ap.add_argument('--precision', choices=['bf16', 'fp32', 'bf16x3', 'bf16x6'], default='bf16')
    ap.add_argument('--impl', choices=['ours', 'reference', 'cudnn'], default='ours')
    ap.add_argument('--profile-step', dest='profile_step', action='store_true', help='warm up, run ONE step, exit (for ncu)')
    ap.add_argument('--no-graph', dest='graph', action='store_false', help='launch kernels one by one instead of replaying a CUDA graph')
    ap.add_argument('--no-multistream', dest='multistream', action='store_false', help='issue every kernel on one stream')
    ap.add_argument('--no-cpu', dest='no_cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--no-also', dest='no_also', action='store_true', help='skip the other BASELINE configs in the default run')
    ap.add_argument('--no-census', dest='no_census', action='store_true', help='skip the per-kernel timing / roofline leg')
    ap.add_argument('--no-gpu-baseline', dest='no_gpu_baseline', action='store_true', help='skip the torch eager + cuDNN leg')
    args = ap.parse_args()
    if args.impl == 'reference':
        if args.batch <= 0:
            args.batch = {'train': 8, 'infer': 8, 'radarnet': 16}[args.mode]
        run_reference(args)
    elif args.impl == 'cudnn':
        if args.batch <= 0:
            args.batch = 8
        run_cudnn(args)
    else:
        if args.warmup < 3 and not args.profile_step:
            args.warmup = 3
        run_ours(args)


if __name__ == '__main__':
    main()
