# round-2 ncu captures (run on the GPU box through gpurun; reports are exported to CSV there: the .ncu-rep files exceed what comes back)
export S4F_NO_GRAPH=1
# the launch list of the TIMED region only (cudaProfilerStart / Stop around it), steady state: warm-up 5, 2 timed outer iterations
S4F_PROFILE_TIMED=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_launches_gamg64.csv python bench.py --steps 2 --warmup 5 --no-cpu-baseline --no-parity > gpurun_out/r2_ncu_bench.log 2>&1
export S4F_TIME_WARMUP=0
ncu --set full --clock-control none -k regex:'k_amg|k_kc|k_amul|k_source|k_grad|k_pcg|k_law|k_tl' -c 48 -f -o /tmp/prof_r2 python profiles/prof_kernels.py 800,100,100 cantilever GAMG > gpurun_out/r2_ncu_full.log 2>&1
ncu -i /tmp/prof_r2.ncu-rep --page raw --csv > gpurun_out/r2_ncu_raw.csv 2>> gpurun_out/r2_ncu_full.log
ncu --set full --clock-control none -k regex:'k_source_g|k_law_mises|k_mises_max_be|k_tl_flux|k_grad' -c 8 -f -o /tmp/prof_r2_notched python profiles/prof_kernels.py 800,100,100 notched_bar GAMG > gpurun_out/r2_ncu_notched.log 2>&1
ncu -i /tmp/prof_r2_notched.ncu-rep --page raw --csv > gpurun_out/r2_ncu_notched_raw.csv 2>> gpurun_out/r2_ncu_notched.log
tail -n 3 gpurun_out/r2_ncu_bench.log gpurun_out/r2_ncu_full.log gpurun_out/r2_ncu_notched.log; ls -la gpurun_out/*.csv
