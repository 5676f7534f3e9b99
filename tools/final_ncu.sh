# ncu --set full captures of the kernels profiles/ reports on; summarised ON THE BOX (the reports exceed gpurun's 64 MiB)
set -x
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:frontend_tma_kernel --launch-skip 2 -c 1 -o /tmp/r2_k1t python tools/prof_frontend.py > gpurun_out/final_ncu_k1t.log 2>&1
python tools/ncu_summary.py /tmp/r2_k1t.ncu-rep frontend_tma_kernel gpurun_out/r2_ncu_frontend_tma.csv
ncu --set full --import-source on --clock-control none -k regex:frontend_tmab_kernel --launch-skip 2 -c 1 -o /tmp/r2_k1tb python tools/prof_frontend.py --format u8 > gpurun_out/final_ncu_k1tb.log 2>&1
python tools/ncu_summary.py /tmp/r2_k1tb.ncu-rep frontend_tmab_kernel gpurun_out/r2_ncu_frontend_tmab_u8.csv
ncu --set full --import-source on --clock-control none -k regex:"pilot_kernel|stereo_kernel|rds_block_kernel|discriminator_kernel" --launch-skip 4 -c 4 -o /tmp/r2_chain python tools/prof_step.py --lanes 1 --streams 64 --steps 1 --warmup 1 > gpurun_out/final_ncu_chain.log 2>&1
for k in pilot_kernel stereo_kernel rds_block_kernel discriminator_kernel; do python tools/ncu_summary.py /tmp/r2_chain.ncu-rep $k gpurun_out/r2_ncu_$k.csv; done
ncu --set full --import-source on --clock-control none -k regex:"fx_dc_par_kernel|frontend_exact_kernel" --launch-skip 2 -c 2 -o /tmp/r2_exact python tools/prof_step.py --front-end-mode 2 --lanes 1 --streams 64 --steps 1 --warmup 1 > gpurun_out/final_ncu_exact.log 2>&1
for k in fx_dc_par_kernel frontend_exact_kernel; do python tools/ncu_summary.py /tmp/r2_exact.ncu-rep $k gpurun_out/r2_ncu_$k.csv; done
ls -la gpurun_out
