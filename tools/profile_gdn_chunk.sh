#!/bin/bash
# ncu evidence for the Gated-Delta-Net prefill kernels (one GPU, run ON the GPU box): the per-kernel launch list of a 4-layer
# Qwen3.5-0.8B-width probe at 4096 rows, then one full capture (with source) of each kernel family.
#   tools/profile_gdn_chunk.sh [launches] [full]        summaries: tools/ncu_launch_summary.py, tools/summarize_ncu.py report
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
PARTS=${@:-launches full}
for part in $PARTS; do case $part in
launches)
    timeout 120 ncu --clock-control none --metrics gpu__time_duration.sum -k regex:gdn_ -c 60 --csv --log-file $O/launches_gdn_chunk.csv \
        python tools/gdn_chunk_probe.py chunked 4096 > $O/ncu_gdn_launches.log 2>&1 ;;
full)
    timeout 150 ncu --clock-control none --set full --import-source on -k "regex:gdn_conv4_qkv|gdn_chunk_|gdn_gated_norm" -s 8 -c 5 -o $O/prof_gdn_chunk -f \
        python tools/gdn_chunk_probe.py chunked 4096 > $O/ncu_gdn_full.log 2>&1 ;;
esac; done
ls -la $O | tail -6
