#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s -k "small_magnitude" 2>&1 | tail -25
