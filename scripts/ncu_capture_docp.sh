#!/bin/bash
# Row f4 evidence: the docp_update sub-object + one `--set full` capture of its kernels,
# summarised on the GPU box.     gpurun -- bash scripts/ncu_capture_docp.sh [tag]
TAG=${1:-v1}
mkdir -p /tmp/ncu gpurun_out
python scripts/prof_docp.py bench > gpurun_out/r02_docp_update_$TAG.json 2> gpurun_out/r02_docp_update_$TAG.err
HQPDOCP_XCOPY=0 python scripts/prof_docp.py bench > gpurun_out/r02_docp_update_${TAG}_noxcopy.json 2>> gpurun_out/r02_docp_update_$TAG.err
ncu --set full --clock-control none --import-source on --kernel-name regex:"docp_" --launch-skip 9 --launch-count 9 \
    -o /tmp/ncu/docp python scripts/prof_docp.py ncu > /tmp/ncu/docp.log 2>&1
python scripts/summarize_ncu.py full /tmp/ncu/docp.ncu-rep \
    "Round 2, ncu --set full of row f4's kernels (Hqp_Docp::update on the device) at config 2's shape, build $TAG" \
    > gpurun_out/r02_ncu_full_docp_$TAG.md
python scripts/ncu_lines.py /tmp/ncu/docp.ncu-rep docp_stage 25 > gpurun_out/r02_ncu_lines_docp_$TAG.txt 2>&1
tail -n 3 /tmp/ncu/docp.log
python - <<PY
import json
for f in ("gpurun_out/r02_docp_update_$TAG.json", "gpurun_out/r02_docp_update_${TAG}_noxcopy.json"):
    d = json.load(open(f))
    print(f, {k: {n: round(v[n]["ms_device"], 4) for n in ("update_ad", "update_fd", "update_fbd")} for k, v in d.items()})
PY
