"""Command-line entry with the reference's flags (reference: src/run_fusionnet.py:8-110):
python run_fusionnet.py --restore_path ckpt.pth --image_path ... (or 'synthetic') --output_dirpath out; add --precision bf16
for the fast mode."""
import argparse

from fusionnet_main import run

parser = argparse.ArgumentParser()
parser.add_argument('--restore_path', type=str, required=True)
parser.add_argument('--image_path', type=str, required=True)
parser.add_argument('--depth_path', type=str, required=True)
parser.add_argument('--response_path', type=str, required=True)
parser.add_argument('--ground_truth_path', type=str, default=None)
parser.add_argument('--input_channels_image', type=int, default=3)
parser.add_argument('--input_channels_depth', type=int, default=2)
parser.add_argument('--normalized_image_range', nargs='+', type=float, default=[0, 1])
parser.add_argument('--encoder_type', nargs='+', type=str, default=['fusionnet18', 'batch_norm'])
parser.add_argument('--n_filters_encoder_image', nargs='+', type=int, default=[32, 64, 128, 256, 256, 256])
parser.add_argument('--n_filters_encoder_depth', nargs='+', type=int, default=[16, 32, 64, 128, 128, 128])
parser.add_argument('--fusion_type', type=str, default='weight_and_project')
parser.add_argument('--decoder_type', nargs='+', type=str, default=['multiscale', 'batch_norm'])
parser.add_argument('--n_filters_decoder', nargs='+', type=int, default=[256, 256, 128, 64, 64, 32])
parser.add_argument('--n_resolutions_decoder', type=int, default=1)
parser.add_argument('--min_predict_depth', type=float, default=1.0)
parser.add_argument('--max_predict_depth', type=float, default=100.0)
parser.add_argument('--weight_initializer', type=str, default='kaiming_uniform')
parser.add_argument('--activation_func', type=str, default='leaky_relu')
parser.add_argument('--output_dirpath', type=str, required=True)
parser.add_argument('--save_outputs', action='store_true')
parser.add_argument('--keep_input_filenames', action='store_true')
parser.add_argument('--verbose', action='store_true')
parser.add_argument('--min_evaluate_depth', type=float, default=0.0)
parser.add_argument('--max_evaluate_depth', type=float, default=100.0)
parser.add_argument('--precision', type=str, default='fp32', choices=['fp32', 'bf16', 'bf16x3', 'bf16x6'],
                    help='B200 path: fp32 parity mode (default) or the bf16 tensor-core mode')

if __name__ == '__main__':
    args = parser.parse_args()
    run(restore_path=args.restore_path, image_path=args.image_path, depth_path=args.depth_path,
        response_path=args.response_path, ground_truth_path=args.ground_truth_path,
        input_channels_image=args.input_channels_image, input_channels_depth=args.input_channels_depth,
        normalized_image_range=args.normalized_image_range, encoder_type=args.encoder_type,
        n_filters_encoder_image=args.n_filters_encoder_image, n_filters_encoder_depth=args.n_filters_encoder_depth,
        fusion_type=args.fusion_type, decoder_type=args.decoder_type, n_filters_decoder=args.n_filters_decoder,
        n_resolutions_decoder=args.n_resolutions_decoder, min_predict_depth=args.min_predict_depth,
        max_predict_depth=args.max_predict_depth, weight_initializer=args.weight_initializer,
        activation_func=args.activation_func, output_dirpath=args.output_dirpath, save_outputs=args.save_outputs,
        keep_input_filenames=args.keep_input_filenames, verbose=args.verbose,
        min_evaluate_depth=args.min_evaluate_depth, max_evaluate_depth=args.max_evaluate_depth, precision=args.precision)
