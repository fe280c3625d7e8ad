#!/bin/bash
# Round-2 (second half) ncu evidence of the large-block path, summarised ON the GPU box:
#   gpurun -- bash scripts/ncu_capture_c4.sh
export HQPCU_GRAPHS=0
mkdir -p /tmp/ncu gpurun_out
# (DMMA runs on the tensor pipe: sm__pipe_fp64_cycles_active does not see it)
DMMA=sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor_subpipe_dmma.sum
FULL="ncu --set full --metrics $DMMA --clock-control none --import-source on"
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/r02_launches_c4.csv python scripts/prof_unit.py c4 1 > /tmp/ncu/c4l.log 2>&1
python scripts/summarize_ncu.py launches gpurun_out/r02_launches_c4.csv \
    "Round 2 (v2), launch list of one unit at nx=200 nu=50 K=296 (C4 stage shape)" > gpurun_out/r02_launches_c4.md
$FULL --kernel-name regex:"seg_element|seg_riccati|elem_hs_kernel" --launch-skip 0 --launch-count 8 \
    -o /tmp/ncu/c4 python scripts/prof_unit.py c4 1 > /tmp/ncu/c4.log 2>&1
python scripts/summarize_ncu.py full /tmp/ncu/c4.ncu-rep \
    "Round 2 (v2: CTA-cooperative GEMM ring, blocked elimination), ncu --set full of the large-block kernels at nx=200 nu=50 (C4 shape, K=296)" \
    > gpurun_out/r02_ncu_full_c4_v2.md
python scripts/ncu_lines.py /tmp/ncu/c4.ncu-rep seg_riccati 25 > gpurun_out/r02_ncu_lines_c4_k3_v2.txt 2>&1
python scripts/ncu_lines.py /tmp/ncu/c4.ncu-rep elem_hs 25 > gpurun_out/r02_ncu_lines_c4_hs_v2.txt 2>&1
tail -n 2 /tmp/ncu/*.log
$FULL --kernel-name regex:k_gemm --launch-skip 7 --launch-count 1 -o /tmp/ncu/gemm ./scratch_bin/mb_biggemm 0 > /tmp/ncu/gemm.log 2>&1
python scripts/summarize_ncu.py full /tmp/ncu/gemm.ncu-rep \
    "Round 2 (v2), ncu --set full of the large-block GEMM in isolation (scripts/mb/mb_biggemm.cu, 200^3, both operands row-contiguous)" \
    > gpurun_out/r02_ncu_full_mb_biggemm.md
