#!/bin/bash
# Round-2 ncu evidence, summarised ON the GPU box (the .ncu-rep files are too large
# to travel back): launch list of one bench run + `--set full` captures of the
# dominant kernels of every workload.   gpurun -- bash scripts/ncu_capture.sh
export HQPCU_GRAPHS=0
mkdir -p /tmp/ncu gpurun_out
FULL="ncu --set full --clock-control none --import-source on"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r02_launches_c2.csv python bench.py --no-extra --steps 2 --warmup 3 \
    > gpurun_out/r02_ncu_bench.log 2>&1
$FULL --kernel-name regex:"elem_hs|seg_element|seg_riccati|elem_terminal" --launch-skip 22 --launch-count 11 \
    -o /tmp/ncu/c2_factor python scripts/prof_unit.py c2 3 > /tmp/ncu/c2.log 2>&1
python scripts/summarize_ncu.py full /tmp/ncu/c2_factor.ncu-rep \
    "Round 2, ncu --set full of the factor kernels (K1, suffix-scan tree, K3) of one unit at C2" \
    > gpurun_out/r02_ncu_full_c2_factor.md
$FULL --kernel-name regex:"solve_" --launch-skip 52 --launch-count 13 \
    -o /tmp/ncu/c2_solve python scripts/prof_unit.py c2 3 > /tmp/ncu/c2s.log 2>&1
python scripts/summarize_ncu.py full /tmp/ncu/c2_solve.ncu-rep \
    "Round 2, ncu --set full of the solve kernels of one step at C2" > gpurun_out/r02_ncu_full_c2_solve.md
$FULL --kernel-name regex:"seg_element|seg_riccati" --launch-skip 2 --launch-count 2 \
    -o /tmp/ncu/c5s python scripts/prof_unit.py c5s 2 > /tmp/ncu/c5s.log 2>&1
python scripts/summarize_ncu.py full /tmp/ncu/c5s.ncu-rep \
    "Round 2, ncu --set full of K1 / K3 at nx=40 nu=10 K=100000 (C5 stage shape)" \
    > gpurun_out/r02_ncu_full_c5s_k1k3.md
$FULL --kernel-name regex:"seg_element|seg_riccati|elem_hs" --launch-skip 0 --launch-count 4 \
    -o /tmp/ncu/c4 python scripts/prof_unit.py c4 1 > /tmp/ncu/c4.log 2>&1
python scripts/summarize_ncu.py full /tmp/ncu/c4.ncu-rep \
    "Round 2, ncu --set full of the large-block kernels at nx=200 nu=50 (C4 shape, K=296)" \
    > gpurun_out/r02_ncu_full_c4.md
$FULL --kernel-name regex:"ips_|residuum" --launch-skip 20 --launch-count 14 \
    -o /tmp/ncu/ips python scripts/prof_unit.py ips > /tmp/ncu/ips.log 2>&1
python scripts/summarize_ncu.py full /tmp/ncu/ips.ncu-rep \
    "Round 2, ncu --set full of the IP vector kernels of one Mehrotra iteration at C2" \
    > gpurun_out/r02_ncu_full_ips.md
python scripts/ncu_lines.py /tmp/ncu/c4.ncu-rep seg_riccati 25 > gpurun_out/r02_ncu_lines_c4_k3.txt 2>&1
python scripts/ncu_lines.py /tmp/ncu/c2_factor.ncu-rep elem_hs 25 > gpurun_out/r02_ncu_lines_c2_hs.txt 2>&1
tail -n 2 /tmp/ncu/*.log
ls -la gpurun_out/r02_*
