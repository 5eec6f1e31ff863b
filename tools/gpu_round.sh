#!/bin/bash
# One GPU-box call: tests, ncu launch list + full capture of the chain kernels, traffic summary, bench.
#   gpurun --timeout 1200 -- 'bash tools/gpu_round.sh r02'
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q > $out/${tag}_tests.log 2>&1; tail -3 $out/${tag}_tests.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 177 -c 177 --csv \
    --log-file $out/${tag}_launches_forward.csv python tools/run_forward.py --iters 2 --batch 64 > $out/${tag}_ncu_l.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tapgemm_chain_kernel -s 26 -c 3 -f -o $out/${tag}_chain_full \
    python tools/run_forward.py --iters 2 --batch 64 > $out/${tag}_ncu_f.log 2>&1
ncu -i $out/${tag}_chain_full.ncu-rep --page details --csv > $out/${tag}_chain_full_details.csv 2>/dev/null
ncu -i $out/${tag}_chain_full.ncu-rep --page raw --csv > $out/${tag}_chain_full_raw.csv 2>/dev/null
python tools/ncu_traffic.py $out/${tag}_chain_full_raw.csv profiles/dominant_traffic.json && cp profiles/dominant_traffic.json $out/${tag}_dominant_traffic.json
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; cat $out/${tag}_bench.json | cut -c1-400
