"""
FusedAdam: torch.optim.Adam semantics (weight_decay 0, amsgrad off -- what the reference builds at
src/fusionnet_main.py:307-312) as ONE kernel over flat buffers (rcfd_adam_step).

Parameters are re-pointed to views of one flat float32 buffer and ``p.grad`` to views of one flat
gradient buffer; the engine's backward writes gradients straight into those views, the
data-parallel all-reduce runs on the flat gradient buffer without copies, and ``step()`` is a
single launch.  Gradients are OVERWRITTEN by every backward (``zero_grad`` only clears the buffer
when asked to), which is what the training loop (zero_grad -> backward -> step) needs.
"""
import torch

from . import engine, ops


class FusedAdam(object):

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        groups = params if (len(params) > 0 and isinstance(params[0], dict)) else [{'params': list(params)}]
        self.params = [p for g in groups for p in g['params']]
        for g in groups:
            if g.get('weight_decay', weight_decay) != 0.0:
                raise ValueError('FusedAdam implements weight_decay = 0 (the reference configuration)')
        self.param_groups = [{'params': self.params, 'lr': lr, 'betas': betas, 'eps': eps, 'weight_decay': 0.0}]
        dev = self.params[0].device
        total = sum(p.numel() for p in self.params)
        self.flat_param = torch.empty(total, device=dev, dtype=torch.float32)
        self.flat_grad = torch.zeros(total, device=dev, dtype=torch.float32)
        self.exp_avg = torch.zeros(total, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(total, device=dev, dtype=torch.float32)
        self.step_count = 0
        off = 0
        for p in self.params:
            n = p.numel()
            self.flat_param[off:off + n].copy_(p.data.reshape(-1))
            p.data = self.flat_param[off:off + n].view(p.shape)
            p.grad = self.flat_grad[off:off + n].view(p.shape)
            p._rcfd_flat = True          # the engine writes this parameter's gradient in place
            off += n

    def zero_grad(self, set_to_none=False):
        """Gradients are overwritten by the next backward; nothing to do (and the views must stay)."""
        return None

    def step(self):
        self.step_count += 1
        g = self.param_groups[0]
        ops.adam_step(self.flat_param, self.flat_grad, self.exp_avg, self.exp_avg_sq, g['lr'], g['betas'][0],
                      g['betas'][1], g['eps'], self.step_count)
        engine.note_params_changed()     # raw-pointer write: torch's version counters did not move

    def state_dict(self):
        return {'state': {'step': self.step_count, 'exp_avg': self.exp_avg, 'exp_avg_sq': self.exp_avg_sq},
                'param_groups': [{k: v for k, v in self.param_groups[0].items() if k != 'params'}]}

    def load_state_dict(self, sd):
        self.step_count = int(sd['state']['step'])
        self.exp_avg.copy_(sd['state']['exp_avg'])
        self.exp_avg_sq.copy_(sd['state']['exp_avg_sq'])
        self.param_groups[0].update(sd['param_groups'][0])
