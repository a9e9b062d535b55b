timeout 300 python -m pytest tests/test_mamba_gpu.py -q -x -k "channels_last" 2>&1 | grep -E "^E|Error|assert" | head -20
