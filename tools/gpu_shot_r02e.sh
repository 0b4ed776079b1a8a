#!/bin/bash
# ncu full captures of the reworked chunk kernels (run ON the GPU box)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 150 ncu --clock-control none --set full --import-source on -k regex:gdn_chunk_ -s 6 -c 3 -o $O/prof_gdn_chunk_v2 -f \
    python tools/gdn_chunk_probe.py chunked 4096 > $O/shote_ncu.log 2>&1; echo "ncu rc $?"
