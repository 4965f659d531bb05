#!/bin/bash
# multi-GPU box: the row-band single-frame mode (codon_group_*) -- tests and timings, each under a short timeout
# (a protocol bug in the cross-GPU ordering would hang, not fail)
timeout 150 python -m pytest tests -m gpu -x -q -k "two_gpus or single_frame_over_gpus" 2>&1 | tail -8
timeout 60 python tools/group_smoke.py f16x3,tf32 2>&1 | tail -4
timeout 240 python tools/run_group.py
