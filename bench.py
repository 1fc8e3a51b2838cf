#!/usr/bin/env python
"""
bench.py -- FusionNet depth-maps/s @352x704 on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--mode train|infer] [--batch B]
                    [--precision bf16|fp32] [--impl ours|reference]

One "step" is one pass of the hot path over one synthetic batch:
  train (default, BASELINE configs[1]): forward + masked-L1 loss + backward + Adam, batch 8 / GPU, bf16;
  infer                               : eval-mode forward, batch 8 / GPU.
`value` is whole-job depth-maps/s with inputs resident in HBM; `e2e` is the same step through
the public API with pinned-host inputs copied H2D and the loss / a depth checksum read back D2H
inside the timed region.  N > 1: launched by torchrun, one rank per GPU, gradients all-reduced
over NCCL (train) or independent replicas (infer); weak scaling.

--impl reference times the reference's own CPU path (the oracle port: same algorithm, fp32,
torch CPU ops on all host threads) on a bounded sample of the same workload.
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
FWD_GFLOP = 57.10            # per depth map, SURVEY.md 8d / BASELINE.md section 2
TRAIN_GFLOP = 170.5
ACT_MB_BF16 = 240.8          # conv activation traffic per depth map, bf16 (fwd)


def read_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d['hbm_gbs'], tc_burst=d['bf16_tflops'], tc_sustained=d['bf16_tflops_sustained'],
                    source='measured (MEASURED_PEAKS.json)')
    return dict(hbm=6650.0, tc_burst=1590.0, tc_sustained=1400.0, source='fallback (B200_PROFILING.md)')


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


# ----------------------------------------------------------------------------- reference arm (CPU oracle port)
def oracle_step_fn(mode, batch, seed=0):
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import fusionnet_oracle as fo
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
        return float(loss)
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
    return best


def time_cpu(mode, batch, steps, warmup):
    pick_cpu_threads()
    step = oracle_step_fn(mode, batch)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return batch / dt, dt


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    batch = 1 if args.mode == 'train' else 2
    value, dt = time_cpu(args.mode, batch, args.steps, args.warmup)
    cores = torch.get_num_threads()
    line = {
        'impl': 'reference', 'metric': 'FusionNet depth-maps/sec @352x704', 'value': value, 'unit': 'depth-maps/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'FusionNet %s step 352x704 (CPU reference path, oracle port of the reference '
                               'algorithm, torch CPU fp32)' % args.mode, 'batch_per_step': batch},
        'cpu_baseline': {'value': value, 'unit': 'depth-maps/s', 'cores': cores, 'kind': 'port',
                         'sample': '%d timed %s steps of batch %d at 352x704 (+%d warm-up), all host threads'
                                   % (args.steps, args.mode, batch, args.warmup)},
        'e2e': {'value': value, 'unit': 'depth-maps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch.distributed as dist
    from rcfd import synth, ops, _lib
    import fusionnet_model
    import net_utils

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    peaks = read_peaks()
    batch = args.batch

    torch.manual_seed(0)
    model = fusionnet_model.FusionNetModel(device=dev, **synth.CANONICAL_FUSIONNET)
    model.set_precision(args.precision)
    model.multistream = args.multistream
    train = args.mode == 'train'
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
    h2d_bytes = sum(t.numel() * 4 for t in (host if train else host[:2]))
    resident = [t.to(dev) for t in host]

    def step(inputs):
        image, depth, gt, lidar = inputs
        if train and args.graph:
            return model.train_step_graphed(image, depth, gt, lidar, opt, 2.0, outlier_removal=outlier)
        if train:
            out = model.forward(image, depth)
            gt_c = outlier.remove_outliers(gt)
            loss, _ = model.compute_loss(image=image, output_depth=out, ground_truth=gt_c, lidar_map=lidar,
                                         loss_func='l1', w_smoothness=0.0, loss_smoothness_kernel_size=-1,
                                         validity_map_loss_smoothness=None, w_lidar_loss=2.0)
            opt.zero_grad()
            loss.backward()
            opt.step()
            return loss
        with torch.no_grad():
            return model.forward_graphed(image, depth) if args.graph else model.forward(image, depth)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms / steps

    # ---- device-resident timing
    for _ in range(args.warmup):
        step(resident)
    if args.profile_step:            # for ncu launch lists: one more step, nothing else
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_push('profile_step')
        step(resident)
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_pop()
        return
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count
    ms = timed(lambda: step(resident), args.steps)
    launches = (_lib.launch_count - l0) // args.steps
    if args.graph:                        # replayed from a CUDA graph: count the kernels captured in it
        launches = getattr(model, 'last_capture_launches', launches)
        if not train:
            l1 = _lib.launch_count
            with torch.no_grad():
                model.forward(resident[0], resident[1])
            launches = _lib.launch_count - l1
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end through the public API: pinned host -> device every step, result read back every step.
    # The graphed entry points stage the host tensors on a copy stream (double buffered), so the copy of step i+1
    # overlaps the compute of step i; the 4-byte result of step i is copied to pinned memory asynchronously and
    # READ by the host while step i+1 runs (the reference reads loss.item() only at checkpoints, fusionnet_main.py:423).
    res_host = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(2)]
    res_ev = [torch.cuda.Event(), torch.cuda.Event()]
    state = {'i': 0, 'last': None}

    def e2e_step():
        if train:           # the graphed step copies straight from the pinned host tensors into its staging buffers
            inputs = host if args.graph else [t.to(dev, non_blocking=True) for t in host]
        else:
            inputs = (host[:2] if args.graph else [t.to(dev, non_blocking=True) for t in host[:2]]) + [None, None]
        r = step(inputs)
        i = state['i']
        res_host[i % 2].copy_((r.detach() if train else r.sum()).reshape(1), non_blocking=True)
        res_ev[i % 2].record()
        if i > 0:                                   # read the previous step's result while this step runs
            res_ev[(i - 1) % 2].synchronize()
            state['last'] = float(res_host[(i - 1) % 2])
        state['i'] = i + 1

    def e2e_drain():
        i = state['i']
        if i > 0:
            res_ev[(i - 1) % 2].synchronize()
            state['last'] = float(res_host[(i - 1) % 2])

    for _ in range(2):
        e2e_step()
    e2e_drain()

    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e2e_drain()                                     # the last result is read inside the timed region too
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t)
    ms_e2e /= args.steps

    # ---- dominant kernel (conv engine) alone on the heaviest layer: decoder deconv0.deconv.conv
    # (64 -> 32, 3x3, nearest 2x up-sampling folded into the loads, output 352 x 704)
    roof = None
    if rank == 0:
        cdt = model.compute_dtype
        x = torch.randn(batch, H // 2, W // 2, 64, device=dev).to(cdt)
        w32 = torch.randn(32, 64, 3, 3, device=dev) * 0.05
        wt = ops.pack_weight(w32, cdt)
        # same call the model makes for this layer: bf16 -> row-streaming sub-pixel engine (needs the phase weights)
        wup = ops.pack_upconv2x_weight(w32, cdt) if cdt == torch.bfloat16 else None
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > L2 (126 MB)
        out = None
        for _ in range(3):
            out = ops.conv2d(x, wt, 32, 3, 1, in_size=(H, W), out=out, engine=model.conv_engine, weight_up2x=wup)
        reps, tot = 5, 0.0
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ops.conv2d(x, wt, 32, 3, 1, in_size=(H, W), out=out, engine=model.conv_engine, weight_up2x=wup)
            b.record()
            torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        k_ms = tot / reps
        esz = 2 if cdt == torch.bfloat16 else 4
        flops = 2.0 * batch * H * W * 32 * 9 * 64
        bytes_ = batch * ((H // 2) * (W // 2) * 64 + H * W * 32) * esz + 32 * 9 * 64 * esz
        tf = flops / (k_ms * 1e-3) / 1e12
        gbs = bytes_ / (k_ms * 1e-3) / 1e9
        if tf / peaks['tc_burst'] >= gbs / peaks['hbm']:
            roof = {'bound': 'tensor', 'achieved': tf, 'peak': peaks['tc_burst'], 'unit': 'TFLOP/s',
                    'frac': tf / peaks['tc_burst']}
        else:
            roof = {'bound': 'hbm', 'achieved': gbs, 'peak': peaks['hbm'], 'unit': 'GB/s', 'frac': gbs / peaks['hbm']}
        traffic = None
        tp = os.path.join(ROOT, 'profiles', 'roofline_kernel_traffic.json')
        if os.path.exists(tp):          # dram__bytes_read.sum + dram__bytes_write.sum of this kernel (ncu --set full, batch 8)
            tj = json.load(open(tp))
            if tj.get('batch') == batch and tj.get('precision') == args.precision:
                traffic = tj.get('dram_bytes_per_launch')
        roof.update({'traffic': traffic, 'executed_gflop': flops / 1e9 * (4.0 / 9.0 if wup is not None else 1.0), 'kernel': 'implicit-GEMM conv, decoder deconv0.deconv (64->32 3x3, fused 2x nearest '
                     'up-sample), batch %d' % batch, 'kernel_ms': k_ms, 'algorithmic_gflop': flops / 1e9,
                     'algorithmic_mb': bytes_ / 1e6, 'achieved_gbs': gbs, 'achieved_tflops': tf, 'peaks': peaks['source']})

    value = world * batch / (ms * 1e-3)
    e2e_value = world * batch / (ms_e2e * 1e-3)
    if rank == 0:
        cpu_batch = 1 if train else 2
        cpu_value, cpu_dt = time_cpu(args.mode, cpu_batch, 1, 1) if (world == 1 and not args.no_cpu) else (None, None)
        gflop = TRAIN_GFLOP if train else FWD_GFLOP
        line = {
            'metric': 'FusionNet depth-maps/sec @352x704', 'value': value, 'unit': 'depth-maps/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'bf16' if args.precision == 'bf16' else 'f32', 'data': 'synthetic',
            'config': {'workload': 'FusionNet %s, 352x704, batch %d per GPU (BASELINE configs[%d])'
                                   % ('training step (fwd + masked-L1 + bwd + Adam)' if train else 'eval forward', batch,
                                      1 if train else 4),
                       'global_batch': world * batch, 'parallelism': 'dp%d' % world,
                       'l2': 'no flush needed: per-step activation working set (>1 GB) exceeds the 126 MB L2',
                       'precision': args.precision, 'cuda_graph': bool(args.graph), 'multistream': bool(args.multistream)},
            'clocks': clocks,
            'e2e': {'value': e2e_value, 'unit': 'depth-maps/s', 'ms_per_step': ms_e2e,
                    'h2d_bytes_per_step': h2d_bytes, 'd2h_bytes_per_step': 4},
            'gpu_launches': launches,
            'roofline': roof,
            'step_roofline': {'tensor_frac': value / world * gflop * 1e9 / (peaks['tc_sustained'] * 1e12),
                              'hbm_frac_fwd_activations': value / world * ACT_MB_BF16 * 1e6 / (peaks['hbm'] * 1e9),
                              'gflop_per_map': gflop, 'peaks': peaks['source']},
            'cpu_baseline': None if cpu_value is None else {
                'value': cpu_value, 'unit': 'depth-maps/s', 'cores': torch.get_num_threads(), 'kind': 'port',
                'sample': '1 timed %s step of batch %d at 352x704 (+1 warm-up), oracle port, thread count picked by a probe'
                          % (args.mode, cpu_batch)},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--mode', choices=['train', 'infer'], default='train')
    ap.add_argument('--batch', type=int, default=8)
    ap.add_argument('--precision', choices=['bf16', 'fp32'], default='bf16')
    ap.add_argument('--impl', choices=['ours', 'reference'], default='ours')
    ap.add_argument('--profile-step', dest='profile_step', action='store_true', help='warm up, run ONE step, exit (for ncu)')
    ap.add_argument('--no-graph', dest='graph', action='store_false', help='launch kernels one by one instead of replaying a CUDA graph')
    ap.add_argument('--no-multistream', dest='multistream', action='store_false', help='issue every kernel on one stream')
    ap.add_argument('--no-cpu', dest='no_cpu', action='store_true', help='skip the cpu_baseline leg')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        if args.warmup < 3 and not args.profile_step:
            args.warmup = 3
        run_ours(args)


if __name__ == '__main__':
    main()
