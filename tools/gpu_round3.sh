#!/bin/bash
# One GPU-box call at the end of round 2: tests, ncu launch lists (bench step, training step), `ncu --set full` of the
# CTA-pair kernels and of the dominant chain kernel, every bench line.
#   gpurun --timeout 2400 -- 'bash tools/gpu_round3.sh r41'
tag=${1:-r41}
out=gpurun_out
mkdir -p $out
if [ -z "$SKIP_TESTS" ]; then timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_tests.log 2>&1; tail -2 $out/${tag}_tests.log; fi
M="--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv"
timeout 300 ncu $M -c 2000 --log-file $out/${tag}_launches_forward.csv python tools/run_forward.py --iters 2 --batch 64 > $out/${tag}_ncu_l.log 2>&1
timeout 600 ncu $M -c 12000 --log-file $out/${tag}_launches_train.csv python tools/run_train.py --steps 2 --batch 32 --eager > $out/${tag}_ncu_t.log 2>&1
F="--set full --clock-control none --kernel-name-base mangled -f"
timeout 300 ncu $F -k regex:pairgemm_kernel -s 3 -c 1 -o $out/${tag}_pairgemm_full python tools/run_forward.py --iters 2 --batch 64 > $out/${tag}_ncu_f1.log 2>&1
timeout 300 ncu $F -k regex:tapgemm_chain_kernel.*Lb0ELb1E -s 10 -c 2 -o $out/${tag}_chainpair_full python tools/run_forward.py --iters 2 --batch 64 > $out/${tag}_ncu_f2.log 2>&1
timeout 300 ncu $F -k regex:tapgemm_chain_kernel.*Li48E.*Lb1ELb0E -s 8 -c 2 -o $out/${tag}_chain_full python tools/run_forward.py --iters 2 --batch 64 > $out/${tag}_ncu_f3.log 2>&1
for n in pairgemm chainpair chain; do
  ncu -i $out/${tag}_${n}_full.ncu-rep --page details --csv > $out/${tag}_${n}_full_details.csv 2>/dev/null
  ncu -i $out/${tag}_${n}_full.ncu-rep --page raw --csv > $out/${tag}_${n}_full_raw.csv 2>/dev/null
done
rm -f $out/${tag}_*_full.ncu-rep      # gpurun brings back <= 64 MiB: the CSV pages are what profiles/ keeps
python tools/ncu_traffic.py $out/${tag}_chain_full_raw.csv profiles/dominant_traffic.json && cp profiles/dominant_traffic.json $out/${tag}_dominant_traffic.json
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; cut -c1-200 $out/${tag}_bench.json
python bench.py --dtype bf16 > $out/${tag}_bench_bf16.json 2> $out/${tag}_bench_bf16.err; cut -c1-200 $out/${tag}_bench_bf16.json
python bench.py --config prenet32 > $out/${tag}_bench_prenet32.json 2> $out/${tag}_bench_prenet32.err; cut -c1-200 $out/${tag}_bench_prenet32.json
python bench.py --config refine3 > $out/${tag}_bench_refine3.json 2> $out/${tag}_bench_refine3.err; cut -c1-200 $out/${tag}_bench_refine3.json
python bench.py --steps 500 --quick > $out/${tag}_bench_steps500.json 2> $out/${tag}_bench_steps500.err; cut -c1-200 $out/${tag}_bench_steps500.json
python bench.py --config train > $out/${tag}_train.json 2> $out/${tag}_train.err; cut -c1-220 $out/${tag}_train.json
python tools/profile_plan.py > $out/${tag}_profile_plan.txt 2>&1; head -3 $out/${tag}_profile_plan.txt
