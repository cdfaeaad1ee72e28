# evidence of the committed build (third session of round 2: raw-coordinate sweep B, shared-space phase 2a, ragged-row-only masking):
# full GPU tests, stress, ncu --set full of the timed kernel, launch list, full bench + reference arm
python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4 > gpurun_out/r2_v13_gputests.txt
timeout 600 python profiles/pipelined_stress.py 120 2>&1 | tail -2 >> gpurun_out/r2_v13_gputests.txt
CMD="python bench.py --scenarios 128 --steps 2 --warmup 3 --no-cpu-baseline --e2e-halfspaces 256"
$CMD > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:pipelined_kernel -s 3 -c 1 -f -o gpurun_out/r2_v13_pipe $CMD > gpurun_out/r2_v13_ncu.log 2>&1
CMD3="python bench.py --dtype f64 --scenarios 64 --steps 2 --warmup 3 --no-cpu-baseline --e2e-halfspaces 128"
$CMD3 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:halfspace_kernel -s 3 -c 1 -f -o gpurun_out/r2_v13_f64 $CMD3 > gpurun_out/r2_v13_ncu_f64.log 2>&1
CMD2="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD2 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_v13_launches.csv $CMD2 > gpurun_out/r2_v13_ncu2.log 2>&1
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_v13_ref.json 2> gpurun_out/r2_v13_ref.err
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_v13_bench.json 2> gpurun_out/r2_v13_bench.err
cat gpurun_out/r2_v13_gputests.txt; python profiles/show_bench.py gpurun_out/r2_v13_bench.json 2>/dev/null | head -3 | cut -c1-900
