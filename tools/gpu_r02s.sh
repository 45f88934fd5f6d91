#!/bin/bash
# eval pre-step on the device: S-level sorter (tcgen05 f16x3 projections, and the FFMA twin), R-level network, RoleOrderer
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s -k "s_ssp or sinkhorn or role_orderer" 2>&1 | tail -12
echo "== VSRDEC_GEMM=simt"
VSRDEC_GEMM=simt timeout 900 python -m pytest tests -m gpu -q -s -k "s_ssp or role_orderer" 2>&1 | tail -8
