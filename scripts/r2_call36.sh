#!/bin/bash
# round 2, GPU call 36: 'resnet' generator: k7 conv cases, upsample+pad, teacher-forced stages, whole network, VanGan step
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "conv3d_fwd_dgrad_wgrad" > $O/r2c36_pytest_conv.log 2>&1; echo "conv rc=$?"; tail -12 $O/r2c36_pytest_conv.log | cut -c1-200
timeout 600 python -m pytest tests/test_gpu_resnet.py -m gpu -q -s > $O/r2c36_pytest_resnet.log 2>&1; echo "resnet rc=$?"; grep -n "resnet stage\|resnet whole\|passed\|failed\|Error\|error\|assert" $O/r2c36_pytest_resnet.log | cut -c1-230 | head -40
