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
        """``torch.optim.Adam.state_dict()`` layout (what the reference stores under 'optimizer_state_dict',
        src/fusionnet_model.py:347-368): per-parameter ``{'step', 'exp_avg', 'exp_avg_sq'}`` keyed by the parameter's
        index, and one param group with the index list and Adam's default flags, so a checkpoint written here loads into
        the reference's ``torch.optim.Adam`` and vice versa.  Moments are copies (not views of the flat buffers).
        torch's Adam only creates state for parameters that received a gradient; parameters whose moments are still
        exactly zero (the never-used projection weights, or every parameter before the first step) are left out the
        same way."""
        state, off = {}, 0
        step = torch.tensor(float(self.step_count))
        for i, p in enumerate(self.params):
            n = p.numel()
            m, v = self.exp_avg[off:off + n], self.exp_avg_sq[off:off + n]
            off += n
            if self.step_count == 0 or not bool(v.any()):
                continue
            state[i] = {'step': step.clone(), 'exp_avg': m.clone().view(p.shape), 'exp_avg_sq': v.clone().view(p.shape)}
        g = self.param_groups[0]
        group = {'lr': g['lr'], 'betas': tuple(g['betas']), 'eps': g['eps'], 'weight_decay': 0.0, 'amsgrad': False,
                 'maximize': False, 'foreach': None, 'capturable': False, 'differentiable': False, 'fused': None,
                 'decoupled_weight_decay': False, 'params': list(range(len(self.params)))}
        return {'state': state, 'param_groups': [group]}

    def load_state_dict(self, sd):
        """Accepts ``torch.optim.Adam.state_dict()`` (a reference checkpoint: per-parameter moments, 'step' an int in
        torch 1.10 or a tensor in current torch) and scatters it into the flat buffers."""
        state = sd['state']
        if 'exp_avg' in state:                      # round-1 private format: flat tensors
            self.step_count = int(state['step'])
            self.exp_avg.copy_(state['exp_avg'])
            self.exp_avg_sq.copy_(state['exp_avg_sq'])
        else:
            groups = sd['param_groups']
            index = [i for g in groups for i in g['params']]
            if len(index) != len(self.params):
                raise ValueError('loaded state dict has %d parameters, the optimizer has %d' % (len(index), len(self.params)))
            self.exp_avg.zero_()
            self.exp_avg_sq.zero_()
            steps, off = [], 0
            offsets = []
            for p in self.params:
                offsets.append(off)
                off += p.numel()
            for pos, key in enumerate(index):
                st = state.get(key, state.get(str(key)))
                if st is None:
                    continue
                p, o = self.params[pos], offsets[pos]
                if st['exp_avg'].numel() != p.numel():
                    raise ValueError('optimizer state of parameter %d has %d elements, expected %d'
                                     % (pos, st['exp_avg'].numel(), p.numel()))
                self.exp_avg[o:o + p.numel()].copy_(st['exp_avg'].reshape(-1))
                self.exp_avg_sq[o:o + p.numel()].copy_(st['exp_avg_sq'].reshape(-1))
                steps.append(int(float(st['step'])))
            if len(set(steps)) > 1:
                raise ValueError('FusedAdam keeps one step count; the loaded state has %s' % sorted(set(steps)))
            self.step_count = steps[0] if steps else 0
        g = sd['param_groups'][0]
        if g.get('weight_decay', 0.0) != 0.0 or g.get('amsgrad', False):
            raise ValueError('FusedAdam implements weight_decay = 0, amsgrad = False (the reference configuration)')
        self.param_groups[0].update({k: g[k] for k in ('lr', 'betas', 'eps') if k in g})
