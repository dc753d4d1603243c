#!/bin/bash
# ncu evidence for one C3 chunk (8192 q-batches, forward + backward) in the default int8 mode with the slice counts the
# probe picks for C3 (6 / 5): (1) launch list with device times, (2) --set full of the LAST forward+backward's own kernels.
# Only CSV exports come back (the .ncu-rep is > 100 MB with sources).
set -x
export MCACQ_CONTRACTION=int8 MCACQ_SLICES=6,5
OUT=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:^(cov_cross|ozaki|posterior_blocks|sample_reduce|scale_inputs|fill_value|reduce_col|unscale_grad|info_summary)" --csv --log-file $OUT/r02_launches_c3_b8192.csv \
    python tools/gpu_fwdbwd_once.py C3 8192 1 > $OUT/r02_launches.log 2>&1
python tools/launch_summary.py $OUT/r02_launches_c3_b8192.csv 16 > $OUT/r02_launches_summary.txt
cat $OUT/r02_launches_summary.txt
TOTAL=$(python - <<'P'
import csv
rows = list(csv.reader(open("gpurun_out/r02_launches_c3_b8192.csv")))
i = next(k for k, r in enumerate(rows) if "Kernel Name" in r)
print(len([r for r in rows[i + 2:] if len(r) > 5]))
P
)
SKIP=$((TOTAL - 16))
echo "launches $TOTAL skip $SKIP"
ncu --set full --clock-control none --import-source on -k "regex:^(cov_cross|ozaki|posterior_blocks|sample_reduce|scale_inputs|fill_value|reduce_col|unscale_grad|info_summary)" -s $SKIP -c 16 -o /tmp/r02_full_c3 \
    python tools/gpu_fwdbwd_once.py C3 8192 1 > $OUT/r02_full.log 2>&1
tail -3 $OUT/r02_full.log
ncu -i /tmp/r02_full_c3.ncu-rep --page raw --csv > $OUT/r02_full_c3_raw.csv
ncu -i /tmp/r02_full_c3.ncu-rep --page source --csv -k regex:ozaki 2>/dev/null | gzip > $OUT/r02_full_c3_ozaki_source.csv.gz
ncu -i /tmp/r02_full_c3.ncu-rep --page details --csv > $OUT/r02_full_c3_details.csv
ls -la /tmp/r02_full_c3.ncu-rep $OUT/
