for rep in 1 2; do
 echo "== pipelined"; python profiles/sustained_power.py 4 2>&1 | tail -1
 echo "== halfspace_kernel (DRCVAR_NO_PIPELINE=1)"; DRCVAR_NO_PIPELINE=1 python profiles/sustained_power.py 4 2>&1 | tail -1
done
