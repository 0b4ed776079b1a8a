#!/bin/bash
# One bounded GPU session for the chunkwise Gated-Delta-Net work (run ON the GPU box): parity of the new kernels first, then the
# config-2 timings in both modes, the per-kernel launch list, the Qwen3.5 / hybrid tests, and -- time permitting -- the rest of the
# GPU suite and the headline bench.  Every stage has its own timeout and writes under gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
date +%s > $O/shot_t0
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > $O/shot_clocks.txt 2>&1
timeout 240 python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "gdn_chunkwise" > $O/shot_t_chunk.log 2>&1; echo "chunk tests rc $?" | tee $O/shot_rc.txt
timeout 80 python tools/gdn_chunk_probe.py chunked 4096 --spans > $O/shot_probe_chunked.json 2> $O/shot_probe_chunked.err; echo "probe chunked rc $?" | tee -a $O/shot_rc.txt
timeout 80 python tools/gdn_chunk_probe.py sequential 4096 --spans > $O/shot_probe_seq.json 2> $O/shot_probe_seq.err; echo "probe seq rc $?" | tee -a $O/shot_rc.txt
CRANE_B200_GDN_PREP=2 timeout 80 python tools/gdn_chunk_probe.py chunked 4096 --spans > $O/shot_probe_chunked_prep2.json 2> $O/shot_probe_chunked_prep2.err; echo "probe chunked prep2 rc $?" | tee -a $O/shot_rc.txt
timeout 120 ncu --clock-control none --metrics gpu__time_duration.sum -k regex:gdn_ -c 60 --csv --log-file $O/shot_launches_gdn_chunk.csv \
    python tools/gdn_chunk_probe.py chunked 4096 > $O/shot_ncu_probe.log 2>&1; echo "ncu launches rc $?" | tee -a $O/shot_rc.txt
CRANE_B200_GDN=chunked timeout 100 python tools/bench_configs.py c2 > $O/shot_c2_chunked.json 2> $O/shot_c2_chunked.err; echo "c2 chunked rc $?" | tee -a $O/shot_rc.txt
CRANE_B200_GDN=sequential timeout 100 python tools/bench_configs.py c2 > $O/shot_c2_seq.json 2> $O/shot_c2_seq.err; echo "c2 seq rc $?" | tee -a $O/shot_rc.txt
timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "qwen3_5 or hybrid or gdn" > $O/shot_t_q35.log 2>&1; echo "qwen3.5 tests rc $?" | tee -a $O/shot_rc.txt
timeout 100 ncu --clock-control none --set full --import-source on -k regex:gdn_chunk_state -s 3 -c 1 -o $O/prof_gdn_chunk_state -f \
    python tools/gdn_chunk_probe.py chunked 4096 > $O/shot_ncu_state.log 2>&1; echo "ncu state rc $?" | tee -a $O/shot_rc.txt
timeout 420 python -m pytest tests -q -m gpu -x > $O/shot_t_all.log 2>&1; echo "all gpu tests rc $?" | tee -a $O/shot_rc.txt
timeout 240 python bench.py > $O/shot_bench.json 2> $O/shot_bench.err; echo "bench rc $?" | tee -a $O/shot_rc.txt
date +%s > $O/shot_t1
tail -3 $O/shot_t_chunk.log; cat $O/shot_probe_chunked.json $O/shot_probe_seq.json $O/shot_c2_chunked.json $O/shot_c2_seq.json 2>/dev/null | cut -c1-600
tail -3 $O/shot_t_q35.log; tail -3 $O/shot_t_all.log; cut -c1-400 $O/shot_bench.json
