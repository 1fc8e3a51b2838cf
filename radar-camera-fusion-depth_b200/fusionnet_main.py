"""
FusionNet training driver with the reference's ``train(...)`` keyword surface (reference:
src/fusionnet_main.py:13-474).  What is accelerated is the step body (reference :348-399): H2D,
forward, ground-truth outlier removal, masked L1 loss, backward, Adam -- all on librcfd_b200.so.

Outside this round's scope and handled explicitly:
  * dataset files: ``rcfd.data.make_train_batches`` reads the reference's path lists / PNG files itself (Pillow for the
    inflate), keeps the on-disk sample types (uint8 / uint16) across PCIe and decodes + crops on the device
    (rcfd_decode_crop); ``train_image_path == 'synthetic'`` gives seeded synthetic batches;
  * augmentation: fusionnet_transforms.Transforms (the reference's draws and arithmetic, batched tensor expressions);
  * validation: ``validate`` (the reference's metrics and best-result rule) runs at checkpoints when a validation set
    is given (``val_image_path='synthetic'`` works too); TensorBoard summaries: skipped.
Multi-GPU: launch with ``torchrun``; ``model.data_parallel()`` attaches the NCCL gradient all-reduce.
"""
import os
import time

import numpy as np
import torch

import eval_utils
from fusionnet_model import FusionNetModel
from fusionnet_transforms import Transforms
from net_utils import OutlierRemoval
from rcfd import data as rcfd_data
from rcfd import optim as rcfd_optim
from rcfd import parallel as rcfd_parallel


def log(text, path=None):
    print(text, flush=True)
    if path is not None:
        with open(path, 'a') as f:
            f.write(text + '\n')


def train(train_image_path, train_depth_path, train_response_path, train_ground_truth_path, train_lidar_map_path,
          val_image_path, val_depth_path, val_response_path, val_ground_truth_path,
          batch_size, n_height, n_width,
          input_channels_image, input_channels_depth, normalized_image_range,
          encoder_type, n_filters_encoder_image, n_filters_encoder_depth, fusion_type, decoder_type,
          n_filters_decoder, n_resolutions_decoder, min_predict_depth, max_predict_depth,
          weight_initializer, activation_func,
          learning_rates, learning_schedule, augmentation_probabilities, augmentation_schedule,
          augmentation_random_crop_type, augmentation_random_brightness, augmentation_random_contrast,
          augmentation_random_saturation, augmentation_random_flip_type,
          loss_func, w_smoothness, w_weight_decay, loss_smoothness_kernel_size, w_lidar_loss,
          ground_truth_outlier_removal_kernel_size, ground_truth_outlier_removal_threshold,
          ground_truth_dilation_kernel_size,
          min_evaluate_depth, max_evaluate_depth,
          checkpoint_dirpath, n_step_per_summary, n_step_per_checkpoint, start_step_validation, restore_path,
          device, n_thread, precision='fp32', max_steps=None, use_cuda_graph=True):
    """Same keyword arguments as the reference (all passed by name); ``precision`` / ``max_steps`` / ``use_cuda_graph`` are extras."""
    if not torch.cuda.is_available():
        raise RuntimeError('fusionnet_main.train needs a CUDA device: the B200 path has no CPU fallback')
    assert len(learning_rates) == len(learning_schedule)
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)          # like the reference (:75), the argument is re-derived
    if world > 1 and not torch.distributed.is_initialized():
        torch.distributed.init_process_group('nccl', device_id=device)

    os.makedirs(checkpoint_dirpath, exist_ok=True)
    checkpoint_path = os.path.join(checkpoint_dirpath, 'model-{}.pth')
    log_path = os.path.join(checkpoint_dirpath, 'results.txt') if rank == 0 else None

    batches, n_train_step_per_epoch = rcfd_data.make_train_batches(
        train_image_path, train_depth_path, train_response_path, train_ground_truth_path, train_lidar_map_path,
        batch_size=batch_size, n_height=n_height, n_width=n_width, crop_type=augmentation_random_crop_type,
        n_thread=n_thread, rank=rank, world=world)

    model = FusionNetModel(
        input_channels_image=input_channels_image, input_channels_depth=input_channels_depth, encoder_type=encoder_type,
        n_filters_encoder_image=n_filters_encoder_image, n_filters_encoder_depth=n_filters_encoder_depth,
        fusion_type=fusion_type, decoder_type=decoder_type, n_resolution_decoder=n_resolutions_decoder,
        n_filters_decoder=n_filters_decoder, deconv_type='up', activation_func=activation_func,
        weight_initializer=weight_initializer, min_predict_depth=min_predict_depth,
        max_predict_depth=max_predict_depth, device=device)
    model.set_precision(precision)
    model.train()
    model.data_parallel()

    learning_rate = learning_rates[0]
    if w_weight_decay == 0.0:
        optimizer = rcfd_optim.FusedAdam([{'params': model.parameters(), 'weight_decay': 0.0}], lr=learning_rate)
        rcfd_parallel.use_flat_gradients(model, optimizer)
    else:
        optimizer = torch.optim.Adam([{'params': model.parameters(), 'weight_decay': w_weight_decay}], lr=learning_rate)

    train_step = 0
    if restore_path is not None and restore_path != '':
        train_step, optimizer = model.restore_model(restore_path, optimizer=optimizer)
        for g in optimizer.param_groups:
            g['lr'] = learning_rate                      # the reference resets the LR on resume (:326-327)

    outlier_removal = None
    if ground_truth_outlier_removal_kernel_size > 1 and ground_truth_outlier_removal_threshold > 0:
        outlier_removal = OutlierRemoval(ground_truth_outlier_removal_kernel_size, ground_truth_outlier_removal_threshold)
    if ground_truth_dilation_kernel_size > 1:
        raise NotImplementedError('ground_truth_dilation_kernel_size > 1 is not used by the shipped configs')
    # augmentation + normalisation exactly like the reference (:300-306, :343-345, :361-364), batched tensor expressions
    train_transforms = Transforms(normalized_image_range=normalized_image_range,
                                  random_brightness=augmentation_random_brightness,
                                  random_contrast=augmentation_random_contrast,
                                  random_saturation=augmentation_random_saturation,
                                  random_flip_type=augmentation_random_flip_type)
    augmentation_schedule_pos = 0
    augmentation_probability = augmentation_probabilities[0]
    # validation set (optional) and the reference's best-result bookkeeping (:134-152, :201-207)
    val_dataloader = rcfd_data.make_val_batches(val_image_path, val_depth_path, val_response_path, val_ground_truth_path,
                                                n_height, n_width) if rank == 0 else None
    val_transforms = Transforms(normalized_image_range=normalized_image_range)
    best_results = {'step': -1, 'mae': np.inf, 'rmse': np.inf, 'imae': np.inf, 'irmse': np.inf}

    def run_validation(step, best):
        model.eval()
        with torch.no_grad():
            best = validate(model=model, dataloader=val_dataloader, transforms=val_transforms, step=step,
                            best_results=best, min_evaluate_depth=min_evaluate_depth,
                            max_evaluate_depth=max_evaluate_depth, device=device, summary_writer=None, log_path=log_path)
        model.train()
        return best
    synthetic = train_image_path == 'synthetic'          # synthetic images are already normalised floats in [0, 1)
    if rank == 0:
        log('Training FusionNet on {} GPU(s), {} steps/epoch, batch {} per GPU, precision {}'.format(
            world, n_train_step_per_epoch, batch_size, precision), log_path)

    # the canonical configuration (l1 + lidar term, no smoothness, FusedAdam) runs through the graphed step
    graphed_step = (use_cuda_graph and loss_func == 'l1' and w_lidar_loss > 0.0 and not w_smoothness > 0.0
                    and isinstance(optimizer, rcfd_optim.FusedAdam))
    learning_schedule_pos = 0
    time_start = time.time()
    n_total = learning_schedule[-1] * n_train_step_per_epoch
    for epoch in range(1, learning_schedule[-1] + 1):
        if epoch > learning_schedule[learning_schedule_pos]:
            learning_schedule_pos += 1
            learning_rate = learning_rates[learning_schedule_pos]
            for g in optimizer.param_groups:
                g['lr'] = learning_rate
        if -1 not in augmentation_schedule and epoch > augmentation_schedule[augmentation_schedule_pos]:
            augmentation_schedule_pos += 1
            augmentation_probability = augmentation_probabilities[augmentation_schedule_pos]
        for batch in batches(epoch):
            train_step += 1
            if getattr(batches, 'raw', False):
                # on-disk sample types (uint8 image, uint16 maps) from pinned memory: value codec, layout and crop on the device
                image, input_depth, input_response, ground_truth, lidar_map = rcfd_data.decode_fusionnet_batch(
                    batch, device, shape=(n_height, n_width))
            else:
                image, input_depth, input_response, ground_truth, lidar_map = [t.to(device, non_blocking=True) for t in batch]
            if not synthetic or augmentation_probability > 0:
                # the reference always goes through Transforms.transform ([0, 255] images in, normalised range out)
                source = (image * 255.0).round() if synthetic else image
                [image], [input_depth, input_response, ground_truth, lidar_map] = train_transforms.transform(
                    images_arr=[source], range_maps_arr=[input_depth, input_response, ground_truth, lidar_map],
                    random_transform_probability=augmentation_probability)
            net_input_depth = torch.cat([input_depth, input_response], dim=1)     # reference :366
            if graphed_step:
                # canonical loss: forward + loss + backward replayed from one CUDA graph (same arithmetic)
                loss = model.train_step_graphed(image, net_input_depth, ground_truth, lidar_map, optimizer,
                                                w_lidar_loss, outlier_removal=outlier_removal)
                if rank == 0 and (train_step % n_step_per_checkpoint) == 0:
                    elapsed = (time.time() - time_start) / 3600
                    remain = (n_total - train_step) * elapsed / max(train_step, 1)
                    log('Step={:6}/{}  Loss={:.5f}  Time Elapsed={:.2f}h  Time Remaining={:.2f}h'.format(
                        train_step, n_total, float(loss), elapsed, remain), log_path)
                    if val_dataloader is not None and train_step >= start_step_validation:
                        best_results = run_validation(train_step, best_results)
                    model.save_model(checkpoint_path.format(train_step), train_step, optimizer)
                if max_steps is not None and train_step >= max_steps:
                    break
                continue
            output_depth = model.forward(image=image, input_depth=net_input_depth)
            if outlier_removal is not None:
                ground_truth = outlier_removal.remove_outliers(ground_truth)
            loss, loss_info = model.compute_loss(
                image=image, output_depth=output_depth, ground_truth=ground_truth, lidar_map=lidar_map,
                loss_func=loss_func, w_smoothness=w_smoothness, loss_smoothness_kernel_size=loss_smoothness_kernel_size,
                # the smoothness term only acts where there is no supervision (reference :379-382)
                validity_map_loss_smoothness=torch.where(ground_truth > 0, torch.zeros_like(ground_truth),
                                                         torch.ones_like(ground_truth)) if w_smoothness > 0 else None,
                w_lidar_loss=w_lidar_loss)
            optimizer.zero_grad()
            loss.backward()
            optimizer.step()
            if rank == 0 and (train_step % n_step_per_checkpoint) == 0:
                elapsed = (time.time() - time_start) / 3600
                remain = (n_total - train_step) * elapsed / max(train_step, 1)
                log('Step={:6}/{}  Loss={:.5f}  Time Elapsed={:.2f}h  Time Remaining={:.2f}h'.format(
                    train_step, n_total, float(loss), elapsed, remain), log_path)
                if val_dataloader is not None and train_step >= start_step_validation:
                    best_results = run_validation(train_step, best_results)
                model.save_model(checkpoint_path.format(train_step), train_step, optimizer)
            if max_steps is not None and train_step >= max_steps:
                break
        if max_steps is not None and train_step >= max_steps:
            break
    if rank == 0:
        if val_dataloader is not None:                       # evaluate once more after training (reference :455-468)
            best_results = run_validation(train_step, best_results)
        model.save_model(checkpoint_path.format(train_step), train_step, optimizer)
        log('Finished at step {} ({:.1f} s)'.format(train_step, time.time() - time_start), log_path)
    return model, optimizer, train_step


def run(restore_path, image_path, depth_path, response_path, ground_truth_path,
        # Input settings
        input_channels_image, input_channels_depth, normalized_image_range,
        # Network settings
        encoder_type, n_filters_encoder_image, n_filters_encoder_depth, fusion_type, decoder_type, n_filters_decoder,
        n_resolutions_decoder, min_predict_depth, max_predict_depth,
        # Weight settings
        weight_initializer, activation_func,
        # Output settings
        output_dirpath, save_outputs, keep_input_filenames, verbose=True,
        # Evaluation settings
        min_evaluate_depth=0.0, max_evaluate_depth=100.0,
        # B200 path
        precision='fp32'):
    """Inference / evaluation over a list of frames with the reference's keyword surface (reference
    src/fusionnet_main.py:608-896): restore a checkpoint (the reference's own files load: 'module.' keys), run every frame
    through the graphed eval forward, evaluate MAE / RMSE (mm) and iMAE / iRMSE (1/km) against the ground truth when it
    is given, optionally save image / ground truth / fused depth / radar depth / radar response with the reference's
    16-bit PNG codecs (src/data_utils.py:271-335).  ``'synthetic'`` as image_path evaluates seeded synthetic frames.
    Returns the dict of mean metrics (None without ground truth).  File reading / writing is host work (Pillow) outside
    the hot path; the depth maps come from the sm_100a kernels (precision: 'fp32' parity mode, 'bf16' fast mode)."""
    device = torch.device('cuda')
    os.makedirs(output_dirpath, exist_ok=True)
    log_path = os.path.join(output_dirpath, 'results.txt')
    ground_truth_available = ground_truth_path is not None and (ground_truth_path == 'synthetic' or os.path.exists(ground_truth_path))
    synthetic = image_path == 'synthetic'
    if synthetic:
        samples = rcfd_data.make_val_batches('synthetic', None, None, None, 352, 704, synthetic_samples=4)
        image_paths = ['synthetic_%d' % i for i in range(len(samples))]
        dataloader = samples if ground_truth_available else [s[:3] for s in samples]
    else:
        image_paths = rcfd_data.read_paths(image_path)
        n = len(image_paths)
        other = [rcfd_data.read_paths(pth) for pth in (depth_path, response_path)]
        gt_paths = rcfd_data.read_paths(ground_truth_path) if ground_truth_available else None
        for paths in other + ([gt_paths] if gt_paths is not None else []):
            assert n == len(paths)
        dataloader = rcfd_data.make_val_batches(image_path, depth_path, response_path,
                                                ground_truth_path if ground_truth_available else None, None, None)
    n_sample = len(image_paths)
    transforms = Transforms(normalized_image_range=normalized_image_range)
    out_dirs = {}
    if save_outputs:
        for name in ('image', 'ground_truth', 'output_depth_fusion', 'output_depth_radar', 'output_response_radar'):
            out_dirs[name] = os.path.join(output_dirpath, name)
            os.makedirs(out_dirs[name], exist_ok=True)

    model = FusionNetModel(
        input_channels_image=input_channels_image, input_channels_depth=input_channels_depth, encoder_type=encoder_type,
        n_filters_encoder_image=n_filters_encoder_image, n_filters_encoder_depth=n_filters_encoder_depth,
        fusion_type=fusion_type, decoder_type=decoder_type, n_resolution_decoder=n_resolutions_decoder,
        n_filters_decoder=n_filters_decoder, deconv_type='up', activation_func=activation_func,
        weight_initializer=weight_initializer, min_predict_depth=min_predict_depth, max_predict_depth=max_predict_depth,
        device=device)
    model.set_precision(precision)
    model.eval()
    model.data_parallel()
    step = -1
    if restore_path:
        step, _ = model.restore_model(restore_path)

    log('Evaluation input paths:', log_path)
    for pth in [image_path, depth_path, response_path] + ([ground_truth_path] if ground_truth_available else []):
        log(str(pth), log_path)
    log('', log_path)
    log('Network settings: encoder {} decoder {} fusion {} filters {} / {} / {} resolutions {} depth [{}, {}] precision {}'.format(
        encoder_type, decoder_type, fusion_type, n_filters_encoder_image, n_filters_encoder_depth, n_filters_decoder,
        n_resolutions_decoder, min_predict_depth, max_predict_depth, precision), log_path)
    log('Evaluation settings: min_evaluate_depth={} max_evaluate_depth={} restore_path={}'.format(
        min_evaluate_depth, max_evaluate_depth, restore_path), log_path)

    mae, rmse, imae, irmse = [np.zeros(n_sample) for _ in range(4)]
    with torch.no_grad():
        for idx, data in enumerate(dataloader):
            data = [datum.to(device) for datum in data]
            if ground_truth_available:
                image, depth, response, ground_truth = data
            else:
                image, depth, response = data[:3]
            [image] = transforms.transform(images_arr=[image], random_transform_probability=0.0)
            input_depth = torch.cat([depth, response], dim=1)
            output_depth = model.forward_graphed(image, input_depth)          # one CUDA graph per frame shape
            output_depth_fusion = np.squeeze(output_depth.cpu().numpy())
            if verbose:
                print('Processed {}/{} samples'.format(idx + 1, n_sample), end='\r')
            if ground_truth_available:
                gt = np.squeeze(ground_truth.cpu().numpy())
                mask = np.where(np.logical_and(gt > 0, np.logical_and(gt > min_evaluate_depth, gt < max_evaluate_depth)))
                mae[idx] = eval_utils.mean_abs_err(1000.0 * output_depth_fusion[mask], 1000.0 * gt[mask])
                rmse[idx] = eval_utils.root_mean_sq_err(1000.0 * output_depth_fusion[mask], 1000.0 * gt[mask])
                imae[idx] = eval_utils.inv_mean_abs_err(0.001 * output_depth_fusion[mask], 0.001 * gt[mask])
                irmse[idx] = eval_utils.inv_root_mean_sq_err(0.001 * output_depth_fusion[mask], 0.001 * gt[mask])
            if save_outputs:
                from PIL import Image
                if keep_input_filenames:
                    filename = os.path.splitext(os.path.basename(image_paths[idx]))[0] + '.png'
                else:
                    filename = '{:010d}.png'.format(idx)
                output_image = np.transpose(np.squeeze(image.cpu().numpy()), (1, 2, 0))
                Image.fromarray((255 * output_image).astype(np.uint8)).save(os.path.join(out_dirs['image'], filename))
                rcfd_data.save_png16(output_depth_fusion, os.path.join(out_dirs['output_depth_fusion'], filename),
                                     rcfd_data.DEPTH_MULTIPLIER)
                rcfd_data.save_png16(np.squeeze(depth.cpu().numpy()), os.path.join(out_dirs['output_depth_radar'], filename),
                                     rcfd_data.DEPTH_MULTIPLIER)
                rcfd_data.save_png16(np.squeeze(response.cpu().numpy()),
                                     os.path.join(out_dirs['output_response_radar'], filename), rcfd_data.RESPONSE_MULTIPLIER)
                if ground_truth_available:
                    rcfd_data.save_png16(gt, os.path.join(out_dirs['ground_truth'], filename), rcfd_data.DEPTH_MULTIPLIER)
    if not ground_truth_available:
        return None
    results = {'mae': float(np.mean(mae)), 'rmse': float(np.mean(rmse)), 'imae': float(np.mean(imae)),
               'irmse': float(np.mean(irmse)), 'step': step}
    log_evaluation_results('Evaluation results', results['mae'], results['rmse'], results['imae'], results['irmse'],
                           step=step, log_path=log_path)
    return results


def validate(model, dataloader, transforms, step, best_results, min_evaluate_depth, max_evaluate_depth, device,
             summary_writer=None, n_summary_display=4, n_summary_display_interval=250, log_path=None):
    """Validation pass with the reference's signature, metrics and best-result rule (reference
    src/fusionnet_main.py:476-606): per sample MAE / RMSE in mm and iMAE / iRMSE in 1/km over the pixels with
    min_evaluate_depth < ground truth < max_evaluate_depth, averaged over the samples; ``best_results`` is replaced
    when more than two of the four metrics (rounded to 2 decimals) are at least as good.  TensorBoard summaries are
    out of scope (``summary_writer`` is accepted and ignored).  The caller puts the model in eval mode."""
    n_sample = len(dataloader)
    mae, rmse, imae, irmse = np.zeros(n_sample), np.zeros(n_sample), np.zeros(n_sample), np.zeros(n_sample)
    for idx, inputs in enumerate(dataloader):
        image, depth, response, ground_truth = [in_.to(device) for in_ in inputs]
        [image] = transforms.transform(images_arr=[image], random_transform_probability=0.0)
        input_depth = torch.cat([depth, response], dim=1)
        with torch.no_grad():
            output_depth = model.forward(image=image, input_depth=input_depth)
        output_depth = np.squeeze(output_depth.cpu().numpy())
        ground_truth = np.squeeze(ground_truth.cpu().numpy())
        mask = np.where(np.logical_and(ground_truth > 0, np.logical_and(ground_truth > min_evaluate_depth,
                                                                         ground_truth < max_evaluate_depth)))
        output_depth, ground_truth = output_depth[mask], ground_truth[mask]
        mae[idx] = eval_utils.mean_abs_err(1000.0 * output_depth, 1000.0 * ground_truth)
        rmse[idx] = eval_utils.root_mean_sq_err(1000.0 * output_depth, 1000.0 * ground_truth)
        imae[idx] = eval_utils.inv_mean_abs_err(0.001 * output_depth, 0.001 * ground_truth)
        irmse[idx] = eval_utils.inv_root_mean_sq_err(0.001 * output_depth, 0.001 * ground_truth)
    mae, rmse, imae, irmse = np.mean(mae), np.mean(rmse), np.mean(imae), np.mean(irmse)
    log_evaluation_results('Validation results', mae, rmse, imae, irmse, step=step, log_path=log_path)
    n_improve = sum(int(np.round(new, 2) <= np.round(best_results[key], 2))
                    for key, new in (('mae', mae), ('rmse', rmse), ('imae', imae), ('irmse', irmse)))
    if n_improve > 2:
        best_results['step'] = step
        best_results['mae'], best_results['rmse'] = mae, rmse
        best_results['imae'], best_results['irmse'] = imae, irmse
    log_evaluation_results('Best results', best_results['mae'], best_results['rmse'], best_results['imae'],
                           best_results['irmse'], step=best_results['step'], log_path=log_path)
    return best_results


def log_evaluation_results(title, mae, rmse, imae, irmse, step=-1, log_path=None):
    """Same table as the reference (:1101-1120)."""
    log(title + ':', log_path)
    log('{:>8}  {:>8}  {:>8}  {:>8}  {:>8}'.format('Step', 'MAE', 'RMSE', 'iMAE', 'iRMSE'), log_path)
    log('{:8}  {:8.3f}  {:8.3f}  {:8.3f}  {:8.3f}'.format(step, mae, rmse, imae, irmse), log_path)
