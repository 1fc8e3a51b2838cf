"""
Stage-1 inference entry point with the reference's signature (reference:
src/radarnet_main.py:534-591): edge-pad the image, run RadarNet on every radar point's
column, then the S2 scatter (paste / threshold / max / arg-max / fill) as ONE kernel
instead of K image-sized temporaries.
"""
import torch

from rcfd import ops

# The reference fills depth through an int64 arg-max tensor: depths are
# truncated and values alias point indices.  True reproduces that bit for bit; False gives
# depth = z[argmax] in float32.
REFERENCE_COMPAT = True


def forward(model, image, radar_points, bounding_boxes_list, device=torch.device('cuda'), compat=None):
    compat = REFERENCE_COMPAT if compat is None else compat
    patch_size = model.input_patch_size_image
    pad_size = patch_size[1] // 2
    # torchvision.transforms.functional.pad(image, (pad, 0, pad, 0), padding_mode='edge') (reference :540-543)
    image = torch.nn.functional.pad(image, (pad_size, pad_size, 0, 0), mode='replicate')
    if radar_points.dim() == 3:
        radar_points = torch.squeeze(radar_points, dim=0)
    output_crops = model.forward(image=image, point=radar_points, bounding_boxes=bounding_boxes_list,
                                 return_logits=False)
    height, width = image.shape[-2], image.shape[-1] - 2 * pad_size
    output_depth, output_response = ops.scatter_tiles_argmax(
        output_crops, radar_points.to(device=output_crops.device, dtype=torch.float32), height, width, compat=compat)
    return output_depth, output_response
