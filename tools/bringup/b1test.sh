python tools/gpu_check_precise.py 2>&1 | grep -i "bit-identical\|ALL OK\|vs oracle\|stage block1\|first chunk of 304\|FAIL\|Error\|Traceback" | head -12
