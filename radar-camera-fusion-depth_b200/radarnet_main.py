"""
Stage-1 entry points with the reference's signatures (reference: src/radarnet_main.py).

``forward`` (:534-591): edge-pad the image, run RadarNet on every radar point's column, then the S2
scatter (paste / threshold / max / arg-max / fill) as ONE kernel instead of K image-sized temporaries.

``train`` (:13-532): the reference's keyword surface; the accelerated part is the step body (:320-403):
label / validity construction from the lidar ground truth and the radar depth, forward, validity-weighted
BCE with logits, backward, Adam -- all on librcfd_b200.so.  Dataset file I/O, TensorBoard summaries and the
validation pass are outside the hot path: batches come from
``rcfd.data.make_radarnet_batches`` (seeded synthetic batches when ``train_image_path == 'synthetic'``).
"""
import os
import time

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


def forward_batch(model, images, radar_points, bounding_boxes, device=torch.device('cuda'), compat=None):
    """``forward`` for N frames in ONE pass (the reference's entry point takes one image: src/radarnet_main.py:534-591):
    images N x 3 x H x W, radar_points N x K x 3, bounding_boxes N x K x 4.  The image encoder runs once over the N frames
    and the decoder once over the N * K point columns, then one S2 scatter per frame; same arithmetic, frame by frame
    identical results.  Returns (depth N x 1 x H x W, response N x 1 x H x W) stacked."""
    compat = REFERENCE_COMPAT if compat is None else compat
    patch_size = model.input_patch_size_image
    pad_size = patch_size[1] // 2
    n, k = radar_points.shape[0], radar_points.shape[1]
    padded = torch.nn.functional.pad(images, (pad_size, pad_size, 0, 0), mode='replicate')
    points = radar_points.reshape(n * k, radar_points.shape[2])
    crops = model.forward(image=padded, point=points, bounding_boxes=[bounding_boxes[b] for b in range(n)],
                          return_logits=False)
    height, width = images.shape[-2], images.shape[-1]
    depth, resp = [], []
    pts = radar_points.to(device=crops.device, dtype=torch.float32)
    for b in range(n):
        d, r = ops.scatter_tiles_argmax(crops[b * k:(b + 1) * k], pts[b], height, width, compat=compat)
        depth.append(d)
        resp.append(r)
    return torch.stack(depth), torch.stack(resp)


def make_labels(ground_truth_depth, radar_depth, max_distance_correspondence, set_invalid_to_negative_class):
    """Ground-truth labels and validity map of the training step (reference :349-378): a pixel of a point's crop is a
    positive when its lidar depth is within ``max_distance_correspondence`` of the radar return's depth; pixels without
    lidar are negatives (and masked out of the loss unless ``set_invalid_to_negative_class``)."""
    distance = torch.abs(ground_truth_depth - radar_depth * torch.ones_like(ground_truth_depth))
    label = torch.where(distance < max_distance_correspondence, torch.ones_like(ground_truth_depth),
                        torch.zeros_like(ground_truth_depth))
    label = torch.where(ground_truth_depth > 0, label, torch.zeros_like(label))
    if set_invalid_to_negative_class:
        validity = torch.ones_like(ground_truth_depth)
    else:
        validity = torch.where(ground_truth_depth <= 0, torch.zeros_like(ground_truth_depth),
                               torch.ones_like(ground_truth_depth))
    return label.float(), validity


def train_step(model, optimizer, image, radar_point, bounding_boxes_list, ground_truth_depth, w_positive_class,
               max_distance_correspondence, set_invalid_to_negative_class):
    """One optimisation step (reference :336-399).  image N x 3 x H x W (edge-padded), radar_point N x K x 3 (x in
    padded-image pixels), bounding_boxes_list N x K x 4, ground_truth_depth N x K x 1 x ph x pw (lidar depth of every
    point's crop).  Returns (loss, logits)."""
    radar_point = radar_point.view(radar_point.shape[0] * radar_point.shape[1], radar_point.shape[2])
    radar_depth = radar_point[..., 2].view(radar_point.shape[0], 1, 1, 1)
    ground_truth_depth = ground_truth_depth.view(ground_truth_depth.shape[0] * ground_truth_depth.shape[1],
                                                 ground_truth_depth.shape[2], ground_truth_depth.shape[3],
                                                 ground_truth_depth.shape[4])
    label, validity = make_labels(ground_truth_depth, radar_depth, max_distance_correspondence,
                                  set_invalid_to_negative_class)
    boxes = [bounding_boxes_list[b] for b in range(bounding_boxes_list.shape[0])]
    logits = model.forward(image, radar_point, boxes, return_logits=True)
    loss, _ = model.compute_loss(logits=logits, ground_truth=label, validity_map=validity,
                                 w_positive_class=w_positive_class)
    optimizer.zero_grad()
    loss.backward()
    optimizer.step()
    return loss, logits


def train(train_image_path, train_radar_path, train_ground_truth_path, val_image_path, val_radar_path,
          val_ground_truth_path,
          batch_size, patch_size, total_points_sampled, sample_probability_of_lidar, normalized_image_range,
          encoder_type, n_filters_encoder_image, n_neurons_encoder_depth, decoder_type, n_filters_decoder,
          weight_initializer, activation_func,
          learning_rates, learning_schedule, augmentation_probabilities, augmentation_schedule,
          augmentation_random_brightness, augmentation_random_contrast, augmentation_random_saturation,
          augmentation_random_noise_type, augmentation_random_noise_spread, augmentation_random_flip_type,
          w_weight_decay, w_positive_class, max_distance_correspondence, set_invalid_to_negative_class,
          checkpoint_dirpath, n_step_per_summary, n_step_per_checkpoint, start_step_validation, restore_path,
          min_evaluate_depth=0.0, max_evaluate_depth=100.0, n_thread=10, precision='fp32', max_steps=None,
          n_height=352, n_width=704):
    """Same keyword arguments as the reference (:13-61); ``precision`` / ``max_steps`` / ``n_height`` / ``n_width``
    (synthetic image size) are extras.  Point noise and flips of the stage-1 augmentation (radarnet_transforms) are data
    preparation outside the hot path: only the colour jitter / normalisation of the image is applied."""
    from radarnet_model import RadarNetModel
    from fusionnet_transforms import Transforms
    from rcfd import data as rcfd_data
    from rcfd import optim as rcfd_optim
    from rcfd import parallel as rcfd_parallel
    if not torch.cuda.is_available():
        raise RuntimeError('radarnet_main.train needs a CUDA device: the B200 path has no CPU fallback')
    if 'none' not in augmentation_random_noise_type and -1 not in augmentation_probabilities[:1] and \
            any(p > 0 for p in augmentation_probabilities) and augmentation_random_noise_type != ['none']:
        raise NotImplementedError('point-noise augmentation (radarnet_transforms) is data preparation, not on the B200 path')
    assert len(learning_rates) == len(learning_schedule)
    local_rank, world, rank = (int(os.environ.get(k, d)) for k, d in (('LOCAL_RANK', '0'), ('WORLD_SIZE', '1'), ('RANK', '0')))
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1 and not torch.distributed.is_initialized():
        torch.distributed.init_process_group('nccl', device_id=device)
    os.makedirs(checkpoint_dirpath, exist_ok=True)
    checkpoint_path = os.path.join(checkpoint_dirpath, 'model-{}.pth')
    log_path = os.path.join(checkpoint_dirpath, 'results.txt') if rank == 0 else None

    def log(text):
        print(text, flush=True)
        if log_path is not None:
            with open(log_path, 'a') as f:
                f.write(text + '\n')

    batches, n_step_per_epoch = rcfd_data.make_radarnet_batches(
        train_image_path, train_radar_path, train_ground_truth_path, batch_size=batch_size, patch_size=patch_size,
        total_points_sampled=total_points_sampled, n_height=n_height, n_width=n_width, rank=rank, world=world)
    model = RadarNetModel(input_channels_image=3, input_channels_depth=3, input_patch_size_image=patch_size,
                          encoder_type=encoder_type, n_filters_encoder_image=n_filters_encoder_image,
                          n_neurons_encoder_depth=n_neurons_encoder_depth, decoder_type=decoder_type,
                          n_filters_decoder=n_filters_decoder, weight_initializer=weight_initializer,
                          activation_func=activation_func, device=device)
    model.set_precision(precision)
    model.train()
    model.data_parallel()
    learning_rate = learning_rates[0]
    if w_weight_decay == 0.0:
        optimizer = rcfd_optim.FusedAdam([{'params': model.parameters(), 'weight_decay': 0.0}], lr=learning_rate)
        rcfd_parallel.use_flat_gradients(model, optimizer)
    else:
        optimizer = torch.optim.Adam([{'params': model.parameters(), 'weight_decay': w_weight_decay}], lr=learning_rate)
    step = 0
    if restore_path is not None and restore_path != '':
        step, optimizer = model.restore_model(restore_path, optimizer=optimizer)
        for g in optimizer.param_groups:
            g['lr'] = learning_rate
    transforms = Transforms(normalized_image_range=normalized_image_range,
                            random_brightness=augmentation_random_brightness,
                            random_contrast=augmentation_random_contrast,
                            random_saturation=augmentation_random_saturation)
    synthetic = train_image_path == 'synthetic'
    augmentation_schedule_pos, augmentation_probability = 0, augmentation_probabilities[0]
    learning_schedule_pos = 0
    n_total = learning_schedule[-1] * n_step_per_epoch
    if rank == 0:
        log('Training RadarNet on {} GPU(s), {} steps/epoch, batch {} x {} points per GPU, precision {}'.format(
            world, n_step_per_epoch, batch_size, total_points_sampled, precision))
    time_start = time.time()
    loss = None
    for epoch in range(1, learning_schedule[-1] + 1):
        if epoch > learning_schedule[learning_schedule_pos]:
            learning_schedule_pos += 1
            learning_rate = learning_rates[learning_schedule_pos]
            for g in optimizer.param_groups:
                g['lr'] = learning_rate
        if -1 not in augmentation_schedule and epoch > augmentation_schedule[augmentation_schedule_pos]:
            augmentation_schedule_pos += 1
            augmentation_probability = augmentation_probabilities[augmentation_schedule_pos]
        for image, radar_point, bounding_boxes, ground_truth_depth in batches(epoch):
            step += 1
            image, radar_point, bounding_boxes, ground_truth_depth = [
                t.to(device, non_blocking=True) for t in (image, radar_point, bounding_boxes, ground_truth_depth)]
            if not synthetic or augmentation_probability > 0:
                source = (image * 255.0).round() if synthetic else image
                [image] = transforms.transform(images_arr=[source], random_transform_probability=augmentation_probability)
            loss, _ = train_step(model, optimizer, image, radar_point, bounding_boxes, ground_truth_depth,
                                 w_positive_class, max_distance_correspondence, set_invalid_to_negative_class)
            if rank == 0 and (step % n_step_per_checkpoint) == 0:
                elapsed = (time.time() - time_start) / 3600
                log('Step={:6}/{} Time Elapsed={:.2f}h  Time Remaining={:.2f}h'.format(
                    step, n_total, elapsed, (n_total - step) * elapsed / max(step, 1)))
                log('Loss={:.5f}'.format(float(loss)))
                model.save_model(checkpoint_path.format(step), step, optimizer)
            if max_steps is not None and step >= max_steps:
                break
        if max_steps is not None and step >= max_steps:
            break
    if rank == 0:
        model.save_model(checkpoint_path.format(step), step, optimizer)
        log('Finished at step {} ({:.1f} s)'.format(step, time.time() - time_start))
    return model, optimizer, step
