import os, sys, tempfile
ROOT='/root/repo'
sys.path.insert(0, os.path.join(ROOT,'radar-camera-fusion-depth_b200'))
import torch, fusionnet_main
def run(graph):
    d=tempfile.mkdtemp()
    kw = dict(train_image_path='synthetic', train_depth_path='synthetic', train_response_path='synthetic',
              train_ground_truth_path='synthetic', train_lidar_map_path='synthetic', val_image_path='',
              val_depth_path='', val_response_path='', val_ground_truth_path='', batch_size=2, n_height=64, n_width=96,
              input_channels_image=3, input_channels_depth=2, normalized_image_range=[0, 1],
              encoder_type=['fusionnet18', 'batch_norm'], n_filters_encoder_image=[16, 16, 32, 32, 32, 32],
              n_filters_encoder_depth=[16, 16, 16, 16, 16, 16], fusion_type='weight_and_project',
              decoder_type=['multiscale', 'batch_norm'], n_filters_decoder=[32, 32, 32, 16, 16, 16],
              n_resolutions_decoder=1, min_predict_depth=1.0, max_predict_depth=100.0,
              weight_initializer='kaiming_uniform', activation_func='leaky_relu', learning_rates=[2e-3, 1e-3],
              learning_schedule=[2, 3], augmentation_probabilities=[0.0], augmentation_schedule=[-1],
              augmentation_random_crop_type=['none'], augmentation_random_brightness=[-1, -1],
              augmentation_random_contrast=[-1, -1], augmentation_random_saturation=[-1, -1],
              augmentation_random_flip_type=['none'], loss_func='l1', w_smoothness=0.0, w_weight_decay=0.0,
              loss_smoothness_kernel_size=-1, w_lidar_loss=2.0, ground_truth_outlier_removal_kernel_size=7,
              ground_truth_outlier_removal_threshold=1.5, ground_truth_dilation_kernel_size=-1, min_evaluate_depth=0.0,
              max_evaluate_depth=100.0, checkpoint_dirpath=d, n_step_per_summary=100,
              n_step_per_checkpoint=1, start_step_validation=1000, restore_path='', device='cuda', n_thread=0,
              precision=sys.argv[1], use_cuda_graph=graph)
    fusionnet_main.train(**kw)
    text=open(os.path.join(d,'results.txt')).read()
    return [float(l.split('Loss=')[1].split()[0]) for l in text.splitlines() if 'Loss=' in l]
for graph in (True, False):
    for rep in range(3):
        print('graph', graph, ['%.3f'%x for x in run(graph)], flush=True)
