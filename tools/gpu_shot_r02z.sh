#!/bin/bash
# Final GPU session of the round: the whole GPU suite, the headline bench line, the config-2 bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 170 python -m pytest tests -q -m gpu -x > $O/shotz_t_all.log 2>&1; echo "all gpu tests rc $?" | tee $O/shotz_rc.txt
timeout 110 python bench.py > $O/shotz_bench.json 2> $O/shotz_bench.err; echo "bench rc $?" | tee -a $O/shotz_rc.txt
timeout 80 python bench.py --config qwen3_5_0_8b --steps 3 > $O/shotz_bench_c2.json 2> $O/shotz_bench_c2.err; echo "bench c2 rc $?" | tee -a $O/shotz_rc.txt
tail -3 $O/shotz_t_all.log; cut -c1-300 $O/shotz_bench.json; cut -c1-1500 $O/shotz_bench_c2.json; tail -2 $O/shotz_bench_c2.err
