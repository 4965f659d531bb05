#!/bin/bash
# 1-GPU box: quick regression + timing after a kernel change
timeout 600 python -m pytest tests -m gpu -x -q -k "forward_matches_reference or image_parity or ragged or cluster_kernels or random_shapes" 2>&1 | tail -4
for m in f16x3 bf16 tf32; do timeout 200 python tools/gpu_class_time.py $m 1 10 | grep "conv5x5\|pair\|cac_apply\|total"; done
timeout 200 python tools/gpu_quick_time.py f16x3 1 20
timeout 200 python tools/gpu_quick_time.py bf16 1 30
