#!/bin/bash
# Time every library in noahmp_b200/exp/ on one B200 (tools/tune_libs.py) and bring the table back in gpurun_out/<name>.log.
#   tools/tune_run.sh NAME [extra shell command run on the box afterwards]
cd "$(dirname "$0")/.."
ARGS=$(for f in noahmp_b200/exp/*.so; do n=$(basename $f .so); echo -n "$n=$f "; done)
/usr/local/graft/bin/gpurun --timeout 1200 -- "python tools/tune_libs.py --ncol 4000000 --reps 4 $ARGS > gpurun_out/$1.log 2>&1; tail -40 gpurun_out/$1.log; ${2:-true}" 2>&1 | grep -v "^\[gpurun\] sending"
