#!/bin/bash
# Follow-up GPU session (one per optimisation step of the chunkwise Gated-Delta-Net work, tag = $1): Qwen3.5 / hybrid parity tests,
# probe timings with stage spans, per-kernel launch list, config-2 timing.  Full captures: tools/profile_gdn_chunk.sh.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
TAG=${1:-d}
timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "qwen3_5 or hybrid or gdn" > $O/shot${TAG}_t_q35.log 2>&1; echo "qwen3.5 tests rc $?" | tee $O/shot${TAG}_rc.txt
timeout 80 python tools/gdn_chunk_probe.py chunked 4096 --spans > $O/shot${TAG}_probe_chunked.json 2> $O/shot${TAG}_probe_chunked.err; echo "probe rc $?" | tee -a $O/shot${TAG}_rc.txt
timeout 100 ncu --clock-control none --metrics gpu__time_duration.sum -k regex:gdn_ -c 60 --csv --log-file $O/shot${TAG}_launches_gdn_chunk.csv \
    python tools/gdn_chunk_probe.py chunked 4096 > $O/shot${TAG}_ncu_probe.log 2>&1; echo "ncu launches rc $?" | tee -a $O/shot${TAG}_rc.txt
timeout 100 python tools/bench_configs.py c2 > $O/shot${TAG}_c2.json 2> $O/shot${TAG}_c2.err; echo "c2 rc $?" | tee -a $O/shot${TAG}_rc.txt
tail -4 $O/shot${TAG}_t_q35.log; cat $O/shot${TAG}_probe_chunked.json $O/shot${TAG}_c2.json
