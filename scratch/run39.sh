for cfg in "16 26" "17 25" "14 28" "15 27" "16 28" "18 26"; do
set -- $cfg
CRBM_DUAL_GY_POS=$1 CRBM_DUAL_GY_NEG=$2 python bench.py --only-headline --no-cpu-baseline --no-scan --no-e2e --steps 2000 > gpurun_out/r3b_$1_$2.json 2> gpurun_out/r3b_$1_$2.err
done
