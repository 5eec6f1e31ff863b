#!/bin/bash
# One GPU-box call of round 2: tests, ncu launch lists of the bench step (inference) and of one training step, bench lines.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round2.sh r09'
tag=${1:-r09}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_tests.log 2>&1; tail -3 $out/${tag}_tests.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2000 --csv \
    --log-file $out/${tag}_launches_forward.csv python tools/run_forward.py --iters 2 --batch 64 > $out/${tag}_ncu_l.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 12000 --csv \
    --log-file $out/${tag}_launches_train.csv python tools/run_train.py --steps 2 --batch 32 --eager > $out/${tag}_ncu_t.log 2>&1
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; cut -c1-300 $out/${tag}_bench.json
python bench.py --config train > $out/${tag}_train.json 2> $out/${tag}_train.err; cut -c1-300 $out/${tag}_train.json
