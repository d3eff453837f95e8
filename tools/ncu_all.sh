#!/bin/bash
# ncu captures of the timed kernels of every workload (one GPU; run under gpurun).  Each workload first runs plain and
# must exit 0.  Output: gpurun_out/<tag>_<workload>.ncu-rep  ->  tools/ncu_to_profile.py
tag=${1:-r02}
shift
wls=${@:-cfg5 cfg2 cfg1 cfg4}
OPS=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum
for wl in $wls; do
  python tools/profile_run.py $wl 2 > gpurun_out/${tag}_${wl}_plain.log 2>&1 && \
  ncu --set full --metrics $OPS --clock-control none -k regex:"(ff_map|ff_prep|ff_pair|nf_map|nf_phase|nf_contract|nf_finalize)" \
      -o gpurun_out/${tag}_${wl} -f python tools/profile_run.py $wl 2 > gpurun_out/${tag}_${wl}_ncu.log 2>&1
  rc=$?
  # the reports are 20-30 MB each: keep the raw-metrics page (and the SASS page of the dominant kernel) as CSV only
  ncu -i gpurun_out/${tag}_${wl}.ncu-rep --page raw --csv > gpurun_out/${tag}_${wl}_raw.csv 2>/dev/null
  ncu -i gpurun_out/${tag}_${wl}.ncu-rep --page source --csv --kernel-name regex:"(ff_map|nf_map|nf_contract)" 2>/dev/null | gzip > gpurun_out/${tag}_${wl}_sass.csv.gz
  rm -f gpurun_out/${tag}_${wl}.ncu-rep
  echo "$wl rc=$rc $(tail -1 gpurun_out/${tag}_${wl}_plain.log)"
done
