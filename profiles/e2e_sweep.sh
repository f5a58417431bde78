#!/bin/bash
# e2e host-fed path: slice schedule sweep (ARTP_SLICE_SCHEDULE) + traced calls (GPU timeline + CPU submit times)
mkdir -p gpurun_out; : > gpurun_out/e2e_sweep.log
for s in ${SCHEDULES:-default "0.1,0.2,0.3,0.4" "0.06,0.14,0.2,0.2,0.2,0.2"}; do
  echo "== $s" >> gpurun_out/e2e_sweep.log
  if [ "$s" = default ]; then unset ARTP_SLICE_SCHEDULE; else export ARTP_SLICE_SCHEDULE=$s; fi
  python profiles/e2e_probe.py 2>&1 | grep "slice" >> gpurun_out/e2e_sweep.log
  ARTP_TRACE=1 python profiles/e2e_probe.py 2>&1 | grep "artp trace" | tail -124 | awk "NR%60==1 || NR%60==2" >> gpurun_out/e2e_sweep.log
done
