#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s -k "s_ssp or sinkhorn or role_orderer" 2>&1 | tail -25
