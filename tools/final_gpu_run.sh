#!/bin/bash
# One GPU-box call at the end of a round: full GPU test suite, smoke, the bench lines that go to profiles/, the ncu
# launch lists of the same commands and one `ncu --set full` capture of the kernels changed this round.
# Usage (from the repo root):  gpurun --timeout 900 -- 'bash tools/final_gpu_run.sh <tag>'
tag=${1:-r02c}
out=gpurun_out
mkdir -p $out
timeout 400 python -m pytest tests -m gpu -q > $out/t_${tag}.log 2>&1; echo rc=$? >> $out/t_${tag}.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_${tag}.log 2>&1; echo rc=$? >> $out/smoke_${tag}.log
timeout 300 python bench.py > $out/b_${tag}_n1.log 2>&1
timeout 120 python bench.py --workload cfg2 --steps 50 --warmup 5 > $out/b_${tag}_cfg2.log 2>&1
timeout 120 python bench.py --workload cfg2_sweep --steps 20 --warmup 5 > $out/b_${tag}_sweep30.log 2>&1
timeout 120 python bench.py --workload cfg2_sweep --heads 6 --steps 20 --warmup 5 > $out/b_${tag}_sweep6.log 2>&1
timeout 200 python bench.py --workload cfg5 > $out/b_${tag}_cfg5.log 2>&1
# launch lists (each command has just exited 0 without the profiler)
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_cfg3.csv python bench.py --steps 2 --warmup 3 > $out/ncu_${tag}_cfg3.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/${tag}_launches_sweep.csv python bench.py --workload cfg2_sweep --steps 4 --warmup 3 > $out/ncu_${tag}_sweep.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/${tag}_launches_cfg2.csv python bench.py --workload cfg2 --steps 10 --warmup 3 > $out/ncu_${tag}_cfg2.log 2>&1
# full captures: the sweep's two tensor-core kernels, the fused exact step
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"sweep_dw_update_tc|sweep_logits_tc" --launch-skip 8 -c 2 -o $out/${tag}_sweep_tc -f python bench.py --workload cfg2_sweep --steps 4 --warmup 3 > $out/ncu_${tag}_sweep_full.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:head_step_fused --launch-skip 20 -c 1 -o $out/${tag}_fused -f python bench.py --workload cfg2 --steps 30 --warmup 3 > $out/ncu_${tag}_fused_full.log 2>&1
echo done > $out/${tag}_done
