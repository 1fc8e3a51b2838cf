"""torchrun --nproc-per-node 2 tools/check_ddp.py : the data-parallel step (overlapped in-graph all-reduces of the flat
gradient buffer) leaves on every rank the MEAN of the per-rank gradients, and all ranks take the same Adam step."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'radar-camera-fusion-depth_b200'))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import fusionnet_model  # noqa: E402
import net_utils  # noqa: E402
from rcfd import optim, parallel, synth  # noqa: E402

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
graph = '--no-graph' not in sys.argv
batch = 2
data = [t.to(dev) for t in bench.synthetic_batch(batch, rank)]
outlier = net_utils.OutlierRemoval(7, 1.5)


def make(ddp):
    torch.manual_seed(0)
    m = fusionnet_model.FusionNetModel(device=dev, **synth.CANONICAL_FUSIONNET)
    m.set_precision('bf16')
    m.train()
    if ddp:
        m.data_parallel()
    opt = optim.FusedAdam(m.parameters(), lr=1e-3)
    if ddp:
        parallel.use_flat_gradients(m, opt)
    return m, opt


def step(m, opt, use_graph):
    if use_graph:
        return m.train_step_graphed(data[0], data[1], data[2], data[3], opt, 2.0, outlier_removal=outlier)
    d = m.forward(data[0], data[1])
    loss, _ = m.compute_loss(data[0], d, outlier.remove_outliers(data[2]), data[3], 'l1', 0.0, -1, None, 2.0)
    loss.backward()
    opt.step()
    return loss


# local gradients (no sync), then their mean over ranks
m0, o0 = make(False)
step(m0, o0, False)
local_grad = o0.flat_grad.clone()
gathered = [torch.empty_like(local_grad) for _ in range(world)]
dist.all_gather(gathered, local_grad)
mean = sum(gathered) / world
# the data-parallel step
m1, o1 = make(True)
p_before = o1.flat_param.clone()
step(m1, o1, graph)
torch.cuda.synchronize()
err = float((o1.flat_grad - mean).abs().max() / mean.abs().max())
# every rank must hold identical parameters after the step
ps = [torch.empty_like(o1.flat_param) for _ in range(world)]
dist.all_gather(ps, o1.flat_param)
same = all(torch.equal(ps[0], q) for q in ps)
moved = float((o1.flat_param - p_before).abs().max())
if rank == 0:
    print('graph=%s  max |grad - mean(local grads)| / max |mean| = %.3e   identical params on all ranks: %s   max param move %.2e'
          % (graph, err, same, moved))
    assert err < 2e-2 and same and moved > 0          # bf16 run-to-run noise between the two executions (atomics order)
    print('DDP CHECK OK')
dist.destroy_process_group()
