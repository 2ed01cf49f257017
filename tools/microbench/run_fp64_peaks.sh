#!/bin/bash
# fp64 peaks of this box's B200 with a clock record (the roofline denominator of bench.py: MEASURED_PEAKS.json carries no
# fp64 figure).  Run under gpurun; writes gpurun_out/fp64_peaks_<tag>.txt.
T=${1:-r2}
cd "$(dirname "$0")"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o fp64_peaks.bin fp64_peaks.cu || exit 1
OUT=../../gpurun_out/fp64_peaks_$T.txt
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv,noheader -lms 200 > /tmp/clk_$T.csv &
SMI=$!
./fp64_peaks.bin > $OUT 2>&1
python ../../tools/microbench/dgemm_torch.py >> $OUT 2>&1
kill $SMI
python - "$T" >> $OUT <<'PY'
import sys
rows = [l.strip().split(", ") for l in open(f"/tmp/clk_{sys.argv[1]}.csv") if l.strip()]
sm = sorted(int(r[0].split()[0]) for r in rows)
busy = [r for r in rows if float(r[2].split()[0]) > 300]
print(f"clocks during the run: {len(rows)} samples, sm MHz median {sm[len(sm)//2]} (min {sm[0]}, max {sm[-1]}), clocks.max.sm {rows[0][1]}, "
      f"power max {max(float(r[2].split()[0]) for r in rows):.0f} W; "
      f"hw_slowdown active in {sum('Active' == r[4] for r in rows)}, hw_thermal {sum('Active' == r[5] for r in rows)}, "
      f"sw_thermal {sum('Active' == r[6] for r in rows)}, sw_power_cap {sum('Active' == r[7] for r in rows)} samples")
PY
cat $OUT
