#!/bin/bash
# ncu --set full of the BatchNorm accumulator kernels, the argmax decode and the Gaussian-target kernel at batch-32 / batch-64 shapes
# (stand-alone diag scripts; each ncu command after the identical plain command exited 0).
mkdir -p gpurun_out
python tools/diag_bn_kernels.py --iters 1 --shapes layer3,layer4 > gpurun_out/bn_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"bn_.*acc_kernel" -s 40 -c 24 -o gpurun_out/prof_bn python tools/diag_bn_kernels.py --iters 1 --shapes layer3,layer4 > gpurun_out/ncu_bn.log 2>&1; echo "ncu bn rc=$?"
python tools/diag_decode.py --iters 1 > gpurun_out/dec_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"argmax_partial_kernel|gauss_targets_kernel" -c 10 -o gpurun_out/prof_dec python tools/diag_decode.py --iters 1 > gpurun_out/ncu_dec.log 2>&1; echo "ncu decode rc=$?"
ls -la gpurun_out/prof_bn.ncu-rep gpurun_out/prof_dec.ncu-rep
