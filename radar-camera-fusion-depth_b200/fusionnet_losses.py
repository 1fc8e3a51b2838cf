"""
Loss functions with the reference's names and signatures (reference: src/fusionnet_losses.py).

Only the masked L1 of the canonical training configuration is on the accelerated hot path
(SURVEY.md 8a a9): ``MaskedL1`` is one fused, synchronisation-free CUDA kernel pair
(rcfd_masked_l1_loss) instead of boolean-mask gathers.  The remaining functions are thin
tensor-op formulas kept for API compatibility (SURVEY.md 8f "next").
"""
import torch

from rcfd import ops


class MaskedL1(torch.autograd.Function):
    """mean|out - gt'| over gt' > 0  +  w_lidar * mean|out - lidar| over lidar > 0,
    gt' = gt where lidar <= 0  (reference src/fusionnet_model.py:214-253, :293)."""

    @staticmethod
    def forward(ctx, output_depth, ground_truth, lidar_map, w_lidar):
        loss, dout = ops.masked_l1_loss(output_depth.detach().float(), ground_truth.float(), lidar_map.float(),
                                        w_lidar, want_grad=True)
        ctx.save_for_backward(dout)
        return loss.view(())

    @staticmethod
    def backward(ctx, g):
        (dout,) = ctx.saved_tensors
        return dout * g, None, None, None


def smooth_l1_loss(src, tgt):
    """mean smooth-L1 (reference :4-17)"""
    return torch.nn.functional.smooth_l1_loss(src, tgt, reduction='mean')


def l1_loss(src, tgt):
    """mean |src - tgt| (reference :19-32)"""
    return torch.mean(torch.abs(src - tgt))


def l2_loss(src, tgt):
    """mean (src - tgt)^2 (reference :34-47)"""
    diff = src - tgt
    return torch.mean(diff * diff)


def gradient_yx(T):
    """forward differences along y and x (reference :131-145)"""
    return T[:, :, :-1, :] - T[:, :, 1:, :], T[:, :, :, :-1] - T[:, :, :, 1:]


class _KernelLoss(torch.autograd.Function):
    """A loss whose value and gradient w.r.t. the prediction come out of one fused kernel call."""

    @staticmethod
    def forward(ctx, predict, fn):
        loss, grad = fn(predict.detach().float().contiguous(), predict.requires_grad)
        ctx.save_for_backward(grad)
        return loss.view(())

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None


def smoothness_loss_func(predict, image):
    """edge-aware first-order smoothness (reference :49-74); CUDA tensors: one fused kernel (rcfd_smoothness_loss)"""
    if predict.is_cuda and not image.requires_grad:
        img = image.detach().float().contiguous()
        return _KernelLoss.apply(predict, lambda p, want: ops.smoothness_loss(p, img, want_grad=want))
    p_dy, p_dx = gradient_yx(predict)
    i_dy, i_dx = gradient_yx(image)
    wx = torch.exp(-i_dx.abs().mean(dim=1, keepdim=True))
    wy = torch.exp(-i_dy.abs().mean(dim=1, keepdim=True))
    return (wx * p_dx.abs()).mean() + (wy * p_dy.abs()).mean()


def sobel_filter(filter_size=[1, 1, 3, 3]):
    """generalised Sobel pair (reference :147-161)"""
    kh, kw = filter_size[-2], filter_size[-1]
    gx, gy = torch.ones(filter_size), torch.ones(filter_size)
    gx[..., kw // 2] = 0
    gx[..., kh // 2, kw // 2 - 1] = 2
    gx[..., kh // 2, kw // 2 + 1] = 2
    gx[..., kw // 2:] = -gx[..., kw // 2:]
    gy[..., kh // 2, :] = 0
    gy[..., kh // 2 - 1, kw // 2] = 2
    gy[..., kh // 2 + 1, kw // 2] = 2
    gy[..., kh // 2 + 1:, :] = -gy[..., kh // 2 + 1:, :]
    return gx, gy


def sobel_smoothness_loss_func(predict, image, weights, filter_size=[1, 1, 7, 7]):
    """Sobel edge-aware smoothness (reference :77-125); CUDA tensors: fused kernels (rcfd_sobel_smoothness_loss)"""
    F = torch.nn.functional
    kh, kw = filter_size[-2], filter_size[-1]
    if (predict.is_cuda and not image.requires_grad and not weights.requires_grad and kh % 2 == 1 and kw % 2 == 1
            and 3 <= kh <= 15 and 3 <= kw <= 15 and image.shape[1] == 3):
        img = image.detach().float().contiguous()
        wts = weights.detach().float().expand(predict.shape).contiguous()
        return _KernelLoss.apply(predict, lambda p, want: ops.sobel_smoothness_loss(p, img, wts, kh, kw, want_grad=want))
    predict = F.pad(predict, (kw // 2, kw // 2, kh // 2, kh // 2), mode='replicate')
    gx, gy = [g.to(predict.device) for g in sobel_filter(filter_size)]
    p_dy, p_dx = F.conv2d(predict, gy), F.conv2d(predict, gx)
    gray = (image[:, 0] * 0.30 + image[:, 1] * 0.59 + image[:, 2] * 0.11).unsqueeze(1)
    gray = F.pad(gray, (1, 1, 1, 1), mode='replicate')
    gxi, gyi = [g.to(predict.device) for g in sobel_filter([1, 1, 3, 3])]
    i_dy, i_dx = F.conv2d(gray, gyi), F.conv2d(gray, gxi)
    wx = torch.exp(-i_dx.abs().mean(dim=1, keepdim=True))
    wy = torch.exp(-i_dy.abs().mean(dim=1, keepdim=True))
    return ((weights * wx * p_dx.abs()).mean() + (weights * wy * p_dy.abs()).mean()) / float(kw * kh)
