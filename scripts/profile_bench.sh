#!/bin/bash
# Run on the GPU box (under gpurun): ncu launch list of the bench command + one full capture of the dominant kernel.
set -x
mkdir -p gpurun_out
R=${1:-r01}
# (1) every launch with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sp_encode|xxh3_chain|index_probe|score_route|prep_rows|index_apply|index_insert" -c 400 --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --requests 16384 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${R}_launches_bench.log 2>&1
# (2) full capture of the tokenizer kernel at the bench's full size (one launch)
ncu --set full --clock-control none --import-source on -k regex:sp_encode_kernel -s 2 -c 1 -o gpurun_out/${R}_sp_encode \
    python scripts/bench_encode.py --n 65536 --iters 1 --warmup 1 --check 0 > gpurun_out/${R}_ncu_sp.log 2>&1
# (3) full capture of the hash kernel at full size
ncu --set full --clock-control none --import-source on -k regex:xxh3_chain128 -s 3 -c 1 -o gpurun_out/${R}_xxh3 \
    python scripts/bench_hash.py --iters 1 > gpurun_out/${R}_ncu_xxh3.log 2>&1
tail -2 gpurun_out/${R}_launches_bench.log | cut -c1-300
