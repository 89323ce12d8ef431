# round-2 final sequence on one B200 (gpurun): full GPU test suite, the bench line, the launch list of the timed region,
# and a compute-sanitizer memcheck pass over a cross-section of the parity tests.  Outputs under gpurun_out/.
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2_final_tests.log
python bench.py > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err
S4F_NO_GRAPH=1 S4F_PROFILE_TIMED=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv \
    --log-file gpurun_out/r2_launches_gamg64.csv python bench.py --steps 2 --warmup 5 --no-cpu-baseline --no-parity > gpurun_out/r2_ncu_bench.log 2>&1
timeout 700 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -q -x \
    -k "test_operators or finite_strain_operators or point_cells_least_squares_gradient or uns_total_lagrangian_face or device_built or gamg_coefficient_refresh or device_mesh_motion" \
    > gpurun_out/r2_memcheck.log 2>&1
echo "memcheck exit code $?" >> gpurun_out/r2_memcheck.log
tail -n 4 gpurun_out/r2_final_tests.log gpurun_out/r2_memcheck.log; tail -c 600 gpurun_out/r2_final_bench_n1.json; wc -l gpurun_out/r2_launches_gamg64.csv
