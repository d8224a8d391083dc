# Round-end evidence on one B200: tests, the default bench line, the reference arm, the ncu launch list of the bench
# command and one --set full capture of a generation's kernels.  Outputs land in gpurun_out/.
set -x
timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 500 python bench.py > gpurun_out/r1_n_bench.json 2> gpurun_out/r1_n_bench.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1_n_bench_reference_arm.json 2> gpurun_out/r1_n_ref.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_n_bench_launches_ncu.csv \
    python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r1_n_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:"k_replicate_fused|k_mutate|k_resample|k_infect_bucket|k_bucket_sort|k_mutcount|k_plan_resample" \
    --launch-skip 140 --launch-count 7 -o gpurun_out/r1_n_full -f python profiles/run_k2.py 24 > gpurun_out/r1_n_ncu_full.log 2>&1
tail -2 gpurun_out/r1_n_ncu_full.log
python profiles/show.py gpurun_out/r1_n_bench.json 2>&1 | head -8
