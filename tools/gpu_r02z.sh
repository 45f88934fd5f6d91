#!/bin/bash
# clock64 breakdown of the single-CTA GEMM kernels inside a real 100-caption decode (debug build, VSR_DBG_CLK)
mkdir -p gpurun_out
VSRDEC_GRAPH=0 VSRDEC_LIB=$PWD/vsr-guided-cic_b200/csrc/build/dbg/libvsrdec_dbg.so timeout 300 python tools/stack_probe.py 100 1 > gpurun_out/r02z_dbg.log 2>&1
grep -E "^gemm" gpurun_out/r02z_dbg.log | awk '{k=$2" "$3" "$4" "$6" "$7; ml[k]+=$10; ep[k]+=$14; n[k]++} END {for (k in n) printf "%s n=%d main %.0f cyc epilogue %.0f cyc\n", k, n[k], ml[k]/n[k], ep[k]/n[k]}' | sort
grep -E "^epi" gpurun_out/r02z_dbg.log | awk '{k=$2" "$3; d[k]+=$6; b1[k]+=$8; pr[k]+=$10; b2[k]+=$12; n[k]++} END {for (k in n) printf "epi %s n=%d dump %.0f bar1 %.0f process %.0f bar2 %.0f\n", k, n[k], d[k]/n[k], b1[k]/n[k], pr[k]/n[k], b2[k]/n[k]}' | sort
grep -E "^cell" gpurun_out/r02z_dbg.log | awk '{k=$2; a[k]+=$6; m[k]+=$8; s[k]+=$10; n[k]++} END {for (k in n) printf "cell %s n=%d loads %.0f math %.0f stores %.0f\n", k, n[k], a[k]/n[k], m[k]/n[k], s[k]/n[k]}' | sort
grep -E "^gemm" gpurun_out/r02z_dbg.log | head -3
tail -1 gpurun_out/r02z_dbg.log | cut -c1-300
