# full driver-style bench of the committed build + ncu evidence (summaries are made on the CPU side)
python bench.py > gpurun_out/r2_t41_bench.json 2> gpurun_out/r2_t41_bench.err; tail -c 600 gpurun_out/r2_t41_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_t41_ref.json 2> gpurun_out/r2_t41_ref.err
CMD="python bench.py --scenarios 128 --steps 2 --warmup 3 --no-cpu-baseline --e2e-halfspaces 256"
$CMD > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:pipelined_kernel -s 3 -c 1 -o gpurun_out/r2_v10b_pipe $CMD > gpurun_out/r2_t41_ncu.log 2>&1
CMD2="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD2 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches.csv $CMD2 > gpurun_out/r2_t41_ncu2.log 2>&1
