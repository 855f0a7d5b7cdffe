#!/usr/bin/env bash
# gpurun_out/prof_<tag>_<workload>.ncu-rep  ->  profiles/<tag>/: raw-page CSV, SASS summary, region table per workload,
# ncu_metrics_<tag>.json (what bench.py's frac_executed reads), metrics table, launch list.
# usage: tools/export_profiles.sh <tag>
set -eu
tag="${1:-r2}"
out=profiles/$tag
mkdir -p $out
args=()
for rep in gpurun_out/prof_${tag}_*.ncu-rep; do
    w=$(basename $rep .ncu-rep); w=${w#prof_${tag}_}
    ncu -i $rep --page raw --csv > $out/ncu_full_$w.csv 2>/dev/null
    ncu -i $rep --page source --csv > /tmp/src_$w.csv 2>/dev/null
    python tools/ncu_sass_summary.py /tmp/src_$w.csv > $out/sass_summary_$w.txt
    ncu -i $rep --page source --csv --print-source cuda,sass > /tmp/srcs_$w.csv 2>/dev/null
    python tools/ncu_regions.py /tmp/srcs_$w.csv 40 > $out/regions_$w.txt || true
    [ "$w" = txt ] && continue
    args+=("$w=$rep")
    rays=$(python - "$w" "$tag" <<'PY'
import re, sys
t = open(f"gpurun_out/plain_{sys.argv[2]}_{sys.argv[1]}.log").read().strip().splitlines()[-1]
m = re.search(r"rays (\d+)", t)
if m:
    print(m.group(1))
else:   # older log lines: kernel ms and Mrays/s only
    ms, mr = re.search(r"kernel ms ([0-9.]+) Mrays/s ([0-9.]+)", t).groups()
    print(round(float(ms) * float(mr) * 1e3))
PY
)
    [ -n "$rays" ] && args+=("rays:$w=$rays")
done
python tools/ncu_metrics_table.py profiles/ncu_metrics_${tag}.json "${args[@]}" > $out/metrics_table.md
cp gpurun_out/launches_${tag}_bench.csv $out/launches_bench_quick.csv 2>/dev/null || true
cp gpurun_out/bench_quick_${tag}.json $out/bench_quick.json 2>/dev/null || true
echo exported to $out
