#!/bin/bash
# Round-2 final ncu evidence (run on the GPU box AFTER the same commands exited 0 without ncu): launch list of the bench,
# --set full captures of the headline CD apply, the NS Jacobian apply and the one-launch partitioned apply (loopback slab).
set -x
O=gpurun_out
NCU="ncu --clock-control none"
python bench.py --steps 5 --warmup 3 --no-extra > $O/r2f_bench_plain.log 2>&1 || exit 1
$NCU --metrics gpu__time_duration.sum -c 80 --csv --log-file $O/r2f_launches_bench.csv python bench.py --steps 5 --warmup 3 --no-extra > $O/r2f_ncu_launches.log 2>&1
$NCU --set full --import-source on -k regex:sem_march3 -s 3 -c 2 -f -o $O/r2f_cd_jvp python bench.py --steps 5 --warmup 3 --no-extra > $O/r2f_ncu_cd.log 2>&1
python scratch/prof_modes.py NS 2 > $O/r2f_prof_modes.log 2>&1 || exit 1
$NCU --set full --import-source on -k regex:sem_march3 -c 4 -f -o $O/r2f_ns_jvp python scratch/prof_modes.py NS 2 > $O/r2f_ncu_ns.log 2>&1
$NCU --set full --import-source on -k regex:sem_march3 -s 35 -c 2 -f -o $O/r2f_xch_slab python scratch/loopback_bench.py 128 20 > $O/r2f_ncu_xch.log 2>&1
ls -la $O/*.ncu-rep
