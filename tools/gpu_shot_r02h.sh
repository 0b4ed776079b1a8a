#!/bin/bash
# GPU session: parity + timings after the tensor-core prep phase, plus full ncu captures of the conv and the chunk kernels
cd "$(dirname "$0")/.."
bash tools/gpu_shot_r02d.sh h
O=gpurun_out
timeout 120 ncu --clock-control none --set full --import-source on -k "regex:gdn_conv4_qkv|gdn_chunk_" -s 8 -c 4 -o $O/prof_gdn_v3 -f \
    python tools/gdn_chunk_probe.py chunked 4096 > $O/shoth_ncu.log 2>&1; echo "ncu rc $?"
