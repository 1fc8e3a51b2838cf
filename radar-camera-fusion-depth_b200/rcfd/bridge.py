"""
In-memory stage-1 -> stage-2 bridge.

The reference glues its two stages through files: `setup/setup_dataset_nuscenes_radarnet.py:293-345` runs RadarNet +
the S2 scatter over the dataset and writes `depth_predicted` / `response_predicted` as 16-bit PNGs
(`src/data_utils.py:271-335`), which FusionNet's data loader reads back (`src/datasets.py`).  Here the same chain runs
on the device for one frame: boxes around every radar point (reference :303-316), RadarNet stage-1 + S2
(`radarnet_main.forward`), the value quantisation of the PNG round trip (optional, on by default because that is
what FusionNet was trained on), FusionNet forward -> dense depth.
"""
import torch

import radarnet_main
from . import ops


def boxes_for_points(radar_points, patch_width, height):
    """Shift the points into the edge-padded image and build their column boxes (x1, 0, x2, H)
    (reference setup/setup_dataset_nuscenes_radarnet.py:303-316, src/radarnet_main.py:980-990)."""
    pad = patch_width // 2
    pts = radar_points.clone().float()
    pts[:, 0] = pts[:, 0] + pad
    zeros = torch.zeros_like(pts[:, 0])
    boxes = torch.stack([pts[:, 0] - pad, zeros, pts[:, 0] + pad, zeros + float(height)], dim=1)
    return pts, [boxes]


def radar_to_input_depth(radarnet_model, image, radar_points, quantize_png16=True, compat=None):
    """image 1 x 3 x H x W in [0, 1], radar_points K x 3 (x, y, z) -> FusionNet input_depth 1 x 2 x H x W."""
    if image.shape[0] != 1:
        raise ValueError('the stage-1 entry point of the reference is per image (batch 1)')
    pts, boxes = boxes_for_points(radar_points.to(image.device), radarnet_model.input_patch_size_image[1], image.shape[-2])
    depth, response = radarnet_main.forward(radarnet_model, image, pts, boxes, device=image.device, compat=compat)
    return ops.stage1_to_stage2(depth, response, quantize_png16=quantize_png16)


def image_and_radar_to_depth(radarnet_model, fusionnet_model, image, radar_points, quantize_png16=True, compat=None):
    """One frame end to end: camera image + radar point cloud -> dense metric depth (1 x 1 x H x W)."""
    with torch.no_grad():
        input_depth = radar_to_input_depth(radarnet_model, image, radar_points, quantize_png16, compat)
        return fusionnet_model.forward(image, input_depth), input_depth
