#!/bin/sh
# One ncu --set full capture of the dominant kernel per configuration (C1 .. C5, reference order) + the launch list of
# bench.py.  Run on a B200 under gpurun; the .ncu-rep files land in gpurun_out/ and are summarised on the build box with
# tools/ncu_summary.py / tools/roofline_inputs.py.
set -e
O=gpurun_out
N=67108864
cap() {  # name, kernel regex, args...
    name=$1; regex=$2; shift 2
    python tools/prof_c2.py "$@" > $O/r02_plain_$name.log 2>&1
    ncu --set full --clock-control none --import-source on -k regex:$regex -s 1 -c 1 -o $O/r02_prof_$name python tools/prof_c2.py "$@" > $O/r02_ncu_$name.log 2>&1
    tail -1 $O/r02_plain_$name.log
}
cap c2_exact sample_seatfirst c2 268435456 4
cap c1 sample_kernel c1 $N
cap c3_exact_interp sample_seatfirst c3 $N 4 0
cap c3_exact_jit bmc_jit_sample c3 $N 4 1
cap c4_exact sample_seatfirst c4 $N 4
cap c5_exact sample_seatfirst c5 268435456 4 0
cap ref sample_reforder ref 16777216
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/r02_bench_plain.json 2> $O/r02_bench_plain.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/r02_bench_under_ncu.log 2>&1
