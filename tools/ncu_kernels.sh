#!/bin/bash
# ncu --set full captures of the named kernels on a reduced batch (2000 subjects of the C4 shape).
# usage (under gpurun): bash tools/ncu_kernels.sh TAG "kernel_regex[:skip]" ...
TAG=$1; shift
CMD="python tools/run_config.py nonseparable 100 6 2000 1"
$CMD > gpurun_out/ncu_plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain_$TAG.log; exit 1; }
for spec in "$@"; do
  k=${spec%%:*}; skip=0
  [[ "$spec" == *:* ]] && skip=${spec##*:}
  name=$(echo "${k}_s${skip}" | tr -c 'A-Za-z0-9_\n' '_')
  ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/prof_${TAG}_$name $CMD > gpurun_out/ncu_${TAG}_$name.log 2>&1
  tail -2 gpurun_out/ncu_${TAG}_$name.log
done
