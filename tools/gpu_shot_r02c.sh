#!/bin/bash
# ncu full captures of the chunk-prep and conv kernels (run ON the GPU box)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 120 ncu --clock-control none --set full --import-source on -k regex:gdn_chunk_prep -s 3 -c 1 -o $O/prof_gdn_chunk_prep -f \
    python tools/gdn_chunk_probe.py chunked 4096 > $O/shotc_ncu_prep.log 2>&1; echo "ncu prep rc $?"
timeout 120 ncu --clock-control none --set full --import-source on -k regex:gdn_conv_kernel -s 3 -c 1 -o $O/prof_gdn_conv -f \
    python tools/gdn_chunk_probe.py chunked 4096 > $O/shotc_ncu_conv.log 2>&1; echo "ncu conv rc $?"
