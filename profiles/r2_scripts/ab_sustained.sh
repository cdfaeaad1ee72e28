# sustained (power-capped) A/B: bash profiles/r2_scripts/ab_sustained.sh secs lib1.so lib2.so ...   (each twice, interleaved)
secs=$1; shift
for rep in 1 2; do for l in "$@"; do echo "== $l"; DRCVAR_LIB=$PWD/$l python profiles/sustained_power.py $secs 2>&1 | tail -1; done; done
