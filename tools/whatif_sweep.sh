F="--no-cpu --no-also --no-census --no-gpu-baseline --steps 10"
run() { echo "== $1"; python tools/whatif.py "$2" -- $F 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'])"; }
run baseline none
run no_wgrad rcfd_conv2d_wgrad,rcfd_unpack_conv_wgrad,rcfd_unpack_stem_s2d_wgrad
run no_pack rcfd_pack_conv_weight,rcfd_pack_upconv2x_weight,rcfd_pack_stem_s2d_weight,rcfd_unpack_conv_wgrad,rcfd_unpack_stem_s2d_wgrad
run no_bn_bwd rcfd_bn_act_bwd_reduce,rcfd_bn_act_bwd_apply
run no_bn_fwd rcfd_bn_train_act_fwd
run no_conv rcfd_conv2d_fwd
run no_adam rcfd_adam_step
python -m pytest tests/test_tc_parity_gpu.py -x -q -m gpu -k "train_step" 2>&1 | tail -3 | cut -c1-300
python tools/timeline.py train 8 2>&1 | tail -75 > gpurun_out/r2_timeline_a.txt; head -8 gpurun_out/r2_timeline_a.txt
