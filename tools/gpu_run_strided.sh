#!/bin/bash
timeout 200 python -m pytest tests/test_autograd_fn_gpu.py -q -s -k "strided or conv_stack or convnext_block_fn" 2>&1 | grep -E "^  |passed|failed|Error|assert" | head -30
