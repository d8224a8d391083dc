# Round-end evidence on one B200: tests, the default bench line, the reference arm, the ncu launch list of the bench
# command and one --set full capture of a generation's kernels.  Outputs land in gpurun_out/ (tag = $1, default r2_z).
tag=${1:-r2_z}
set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_ref.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_bench_launches_ncu.csv \
    python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/${tag}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:"k_mutate|k_offspring_sum|k_plan_resample|k_resample_pos|k_infect_sum|k_gc_copy" \
    --launch-skip 82 --launch-count 6 -o gpurun_out/${tag}_full -f python profiles/run_k2.py 24 > gpurun_out/${tag}_ncu_full.log 2>&1
tail -2 gpurun_out/${tag}_ncu_full.log
python profiles/show.py gpurun_out/${tag}_bench.json 2>&1 | head -12
