timeout 600 python -m pytest tests -m gpu -q --timeout 300 -s -k "fused_kernels_match" 2>&1 | grep -E "fused vs|passed|failed|Error" | head
