#!/bin/bash
# One gpurun call: GPU tests, smoke, the bench line, and the ncu evidence of the same command (launch list, per-launch
# traffic of conv_os_kernel, one full capture).  Outputs under gpurun_out/.
O=gpurun_out
mkdir -p $O
timeout 400 python -m pytest tests -m gpu -x -q > $O/f_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/f_rc.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/f_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/f_rc.log
timeout 400 python bench.py > $O/f_bench_n1.json 2> $O/f_bench_n1.err; echo "bench rc=$?" | tee -a $O/f_rc.log
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-scaling-baseline --no-roofline --no-graph"
timeout 200 $B > $O/f_eager.log 2>&1; echo "eager rc=$?" | tee -a $O/f_rc.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 2700 --csv --log-file $O/f_launches.csv $B > $O/f_ncu1.log 2>&1; echo "ncu-list rc=$?" | tee -a $O/f_rc.log
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:conv_os_kernel -s 164 -c 82 --csv --log-file $O/f_traffic.csv $B > $O/f_ncu2.log 2>&1; echo "ncu-traffic rc=$?" | tee -a $O/f_rc.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_os_kernel -s 120 -c 3 -o $O/f_conv_os_full -f $B > $O/f_ncu3.log 2>&1; echo "ncu-full rc=$?" | tee -a $O/f_rc.log
tail -3 $O/f_pytest.log; tail -2 $O/f_smoke.log; head -c 600 $O/f_bench_n1.json
