# evidence of the committed build: ncu --set full of the timed kernel, launch list, full bench + reference arm
CMD="python bench.py --scenarios 128 --steps 2 --warmup 3 --no-cpu-baseline --e2e-halfspaces 256"
$CMD > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:pipelined_kernel -s 3 -c 1 -f -o gpurun_out/r2_final_pipe $CMD > gpurun_out/r2_final_ncu.log 2>&1
CMD3="python bench.py --dtype f64 --scenarios 64 --steps 2 --warmup 3 --no-cpu-baseline --e2e-halfspaces 128"
$CMD3 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:halfspace_kernel -s 3 -c 1 -f -o gpurun_out/r2_final_f64 $CMD3 > gpurun_out/r2_final_ncu_f64.log 2>&1
CMD2="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD2 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_final_launches.csv $CMD2 > gpurun_out/r2_final_ncu2.log 2>&1
python bench.py --impl reference > gpurun_out/r2_final_ref.json 2> gpurun_out/r2_final_ref.err
python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err
