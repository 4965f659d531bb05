#!/bin/bash
# GPU box: f16x3 (split-fp16) mode against the oracle next to fp32, plus device time per frame.
set -x
timeout 300 python tests/checkers/debug_taps.py f16x3 2 45 70
timeout 300 python tests/checkers/debug_taps.py fp32 2 45 70
timeout 300 python tests/checkers/debug_taps.py f16x3 1 160 240
timeout 600 python tests/checkers/debug_taps.py f16x3 1 480 640
timeout 300 python tests/checkers/debug_taps.py f16x3 1 1 1
timeout 300 python tests/checkers/debug_taps.py f16x3 3 200 1
timeout 300 python tools/gpu_quick_time.py f16x3 1 20
timeout 300 python tools/gpu_quick_time.py tf32 1 20
timeout 300 python tools/gpu_quick_time.py bf16 1 20
timeout 300 python tools/gpu_quick_time.py f16x3 4 10
