tag=r02d
out=gpurun_out
timeout 400 python -m pytest tests -m gpu -q > $out/t_${tag}.log 2>&1; echo rc=$? >> $out/t_${tag}.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_${tag}.log 2>&1; echo rc=$? >> $out/smoke_${tag}.log
timeout 120 python bench.py --workload cfg2 --steps 50 --warmup 5 > $out/b_${tag}_cfg2.log 2>&1
timeout 120 python bench.py --workload cfg2_sweep --steps 20 --warmup 5 > $out/b_${tag}_sweep30.log 2>&1
timeout 120 python bench.py --workload cfg2_sweep --heads 6 --steps 20 --warmup 5 > $out/b_${tag}_sweep6.log 2>&1
timeout 200 python bench.py --workload cfg5 > $out/b_${tag}_cfg5.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/${tag}_launches_sweep.csv python bench.py --workload cfg2_sweep --steps 4 --warmup 3 > $out/ncu_${tag}_sweep.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"sweep_dw_update_tc|sweep_logits_tc" --launch-skip 8 -c 2 -o $out/${tag}_sweep_tc -f python bench.py --workload cfg2_sweep --steps 4 --warmup 3 > $out/ncu_${tag}_sweep_full.log 2>&1
echo done > $out/${tag}_done
