export S4F_NO_GRAPH=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches_gamg64.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/r2_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_amg|k_kc|k_amul|k_source|k_grad|k_pcg|k_law|k_tl' -c 230 -f -o gpurun_out/prof_r2 python profiles/prof_kernels.py 800,100,100 cantilever GAMG > gpurun_out/r2_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_source_g|k_law_mises|k_mises_max_be|k_tl_flux|k_grad' -c 24 -f -o gpurun_out/prof_r2_notched python profiles/prof_kernels.py 800,100,100 notched_bar GAMG > gpurun_out/r2_ncu_notched.log 2>&1
tail -3 gpurun_out/r2_ncu_bench.log gpurun_out/r2_ncu_full.log gpurun_out/r2_ncu_notched.log; ls -la gpurun_out/*.ncu-rep
