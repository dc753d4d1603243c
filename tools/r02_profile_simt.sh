#!/bin/bash
# ncu --set full with source counters of the FP64-side kernels of one C3 chunk (timed pass only): per-kernel source pages
set -x
export MCACQ_CONTRACTION=int8 MCACQ_SLICES=6,5 MCACQ_PROFILE_LAST=1
OUT=gpurun_out
K="regex:^(cov_cross|sample_reduce|posterior_blocks)"
ncu --set full --clock-control none --import-source on --profile-from-start off -k "$K" -o /tmp/r02_simt \
    python tools/gpu_fwdbwd_once.py C3 8192 1 > $OUT/r02_simt.log 2>&1
tail -2 $OUT/r02_simt.log
for k in cov_cross_kernel cov_cross_bwd_kernel sample_reduce_fwd_kernel sample_reduce_bwd_kernel posterior_blocks_bwd_kernel; do
  ncu -i /tmp/r02_simt.ncu-rep --page source --csv -k regex:$k 2>/dev/null | gzip > $OUT/r02_simt_${k}_source.csv.gz
done
ncu -i /tmp/r02_simt.ncu-rep --page raw --csv > $OUT/r02_simt_raw.csv
ls -la $OUT/r02_simt* /tmp/r02_simt.ncu-rep
