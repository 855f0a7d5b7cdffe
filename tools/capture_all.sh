#!/usr/bin/env bash
# Run ON THE GPU BOX (through gpurun): one `ncu --set full` capture of the render kernel per workload (third launch,
# after two warm-up launches), the FP32 thread-instruction counters next to it, and the launch list of a short
# bench.py run.  Everything lands in gpurun_out/; tools/export_profiles.sh turns it into profiles/<tag>/.
# usage: [WORKLOADS="a b"] tools/capture_all.sh <tag>     (WORKLOADS: only these render captures)
set -u
tag="${1:-r2}"
FP32=smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum
for w in ${WORKLOADS:-default_1080p_d5 default_4k_d50 synth1024_1080p_d50 synth256_1080p_d10 two_mirrors_1080p_d50 default_500x504_d50}; do
    python tools/profile_run.py $w 3 > gpurun_out/plain_${tag}_$w.log 2>&1 || { echo "plain run of $w failed"; continue; }
    ncu --set full --metrics $FP32 --import-source on --clock-control none -k "regex:render_(grid_)?kernel" -s 2 -c 1 \
        -o gpurun_out/prof_${tag}_$w -f python tools/profile_run.py $w 3 > gpurun_out/ncu_${tag}_$w.log 2>&1
done
[ -n "${WORKLOADS:-}" ] && { echo capture_all done; exit 0; }
python tools/profile_run.py default_1080p_d5 2 --txt > gpurun_out/plain_${tag}_txt.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:txt_fixed_kernel -c 1 -o gpurun_out/prof_${tag}_txt -f \
    python tools/profile_run.py default_1080p_d5 2 --txt > gpurun_out/ncu_${tag}_txt.log 2>&1
python bench.py --steps 2 --warmup 1 --quick --no-cpu-baseline > gpurun_out/bench_quick_${tag}.json 2> gpurun_out/bench_quick_${tag}.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}_bench.csv \
    python bench.py --steps 2 --warmup 1 --quick --no-cpu-baseline > gpurun_out/ncu_${tag}_launches.log 2>&1
echo capture_all done
