#!/bin/bash
# runs one bench step of a -DK2Y_NMS_TRACE build and prints the phase clocks of the dense classes of image 0; parity tests first
timeout 600 python -m pytest tests/test_gpu_detect.py tests/test_gpu_fullsize.py -q -m gpu --timeout 300 -x 2>&1 | grep -v nms-trace | tail -3
timeout 300 python bench.py --profile-step --no-cpu 2>&1 | grep "nms-trace" | sort | grep "chunk 0/" | tail -12
