#!/bin/bash
# Round-2 ncu evidence (one GPU, run ON the GPU box).  Usage: tools/profile_r02.sh [part ...]   parts: launches ll others prefill
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PARTS=${@:-launches ll others}
BENCH="python bench.py --steps 1 --warmup 0 --no-cpu-baseline"
NCU="ncu --clock-control none"
for part in $PARTS; do case $part in
launches)
    # every launch of one request with its device time (cold-cache, serialised: compare SHARES): ViT + prefill kernels, ONE persistent
    # decode launch (256 steps), then the e2e leg's per-token launches
    timeout 900 $NCU --metrics gpu__time_duration.sum -c 900 --csv --log-file gpurun_out/launches_r02.csv $BENCH > gpurun_out/ncu_launches_r02.log 2>&1 ;;
ll)
    # the persistent decode kernel: DRAM traffic of the 256-step launch, then one full-set capture
    timeout 900 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum -k regex:decode_ll_kernel -c 1 --csv \
        --log-file gpurun_out/ll_traffic_r02.csv $BENCH > gpurun_out/ncu_ll_traffic_r02.log 2>&1
    timeout 1200 $NCU --set full --import-source on -k regex:decode_ll_kernel -c 1 -o gpurun_out/prof_ll_r02 -f $BENCH > gpurun_out/ncu_ll_r02.log 2>&1 ;;
others)
    # kernels of the other BASELINE configs (tools/bench_configs.py): quantised GEMV + activation quantiser (c3), GatedDeltaNet scan /
    # decode + D=256 attention (c2), the TTS frame (c4)
    timeout 600 $NCU --set full --import-source on -k regex:qgemv -s 200 -c 4 -o gpurun_out/prof_qgemv_r02 -f python tools/bench_configs.py c3 > gpurun_out/ncu_qgemv_r02.log 2>&1
    timeout 600 $NCU --set full --import-source on -k regex:xquant_kernel -s 100 -c 2 -o gpurun_out/prof_xquant_r02 -f python tools/bench_configs.py c3 > gpurun_out/ncu_xquant_r02.log 2>&1
    timeout 600 $NCU --set full --import-source on -k "regex:gdn_" -s 2 -c 6 -o gpurun_out/prof_gdn_r02 -f python tools/bench_configs.py c2 > gpurun_out/ncu_gdn_r02.log 2>&1
    timeout 600 $NCU --set full --import-source on -k "regex:attn_decode_kernel<256" -s 2 -c 2 -o gpurun_out/prof_attn256_r02 -f python tools/bench_configs.py c2 > gpurun_out/ncu_attn256_r02.log 2>&1
    CRANE_B200_GRAPHS=0 timeout 600 $NCU --metrics gpu__time_duration.sum -c 1500 --csv --log-file gpurun_out/launches_tts_r02.csv python tools/bench_configs.py c4 > gpurun_out/ncu_tts_r02.log 2>&1 ;;
prefill)
    # the prefill kernels after this round's rewrites: persistent stream-K GEMM (text layer: qkv, o, gate/up, down) and the
    # two-group flash attention (ViT + text)
    timeout 600 $NCU --set full --import-source on -k regex:gemm_tc_kernel -s 230 -c 6 -o gpurun_out/prof_gemm_r02 -f $BENCH > gpurun_out/ncu_gemm_r02.log 2>&1
    timeout 600 $NCU --set full --import-source on -k regex:flash_prefill_kernel -s 30 -c 2 -o gpurun_out/prof_flash_r02 -f $BENCH > gpurun_out/ncu_flash_r02.log 2>&1 ;;
esac; done
ls -la gpurun_out | tail -24
