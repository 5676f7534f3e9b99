# ncu --set full captures of the kernels profiles/ reports on; summarised ON THE BOX (the reports exceed gpurun's 64 MiB)
set -x
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:frontend_tma_kernel --launch-skip 2 -c 1 -o /tmp/r2_k1t python tools/prof_frontend.py > gpurun_out/final_ncu_k1t.log 2>&1
python tools/ncu_summary.py /tmp/r2_k1t.ncu-rep frontend_tma_kernel gpurun_out/r2_ncu_frontend_tma.csv
ncu --set full --import-source on --clock-control none -k regex:frontend_tmab_kernel --launch-skip 2 -c 1 -o /tmp/r2_k1tb python tools/prof_frontend.py --format u8 > gpurun_out/final_ncu_k1tb.log 2>&1
python tools/ncu_summary.py /tmp/r2_k1tb.ncu-rep frontend_tmab_kernel gpurun_out/r2_ncu_frontend_tmab_u8.csv
# whole-call launches (time slicing off), after 1.5 s of signal so that the pilot is locked and the PSS loop runs:
# per step the regex matches discriminator, pilot, stereo and three rds_block launches
SDRJFM_FM_SLICE=0 ncu --set full --import-source on --clock-control none -k regex:"pilot_kernel|stereo_kernel|rds_block_kernel|discriminator_kernel" --launch-skip 18 -c 6 -o /tmp/r2_chain python tools/prof_step.py --lanes 1 --streams 64 --steps 1 --warmup 3 > gpurun_out/final_ncu_chain.log 2>&1
for k in pilot_kernel stereo_kernel rds_block_kernel discriminator_kernel; do python tools/ncu_summary.py /tmp/r2_chain.ncu-rep $k gpurun_out/r2_ncu_$k.csv; done
ncu --set full --import-source on --clock-control none -k regex:"fx_dc_par_kernel|frontend_exact_kernel" --launch-skip 4 -c 2 -o /tmp/r2_exact python tools/prof_step.py --front-end-mode 2 --lanes 1 --streams 64 --steps 1 --warmup 2 > gpurun_out/final_ncu_exact.log 2>&1
for k in fx_dc_par_kernel frontend_exact_kernel; do python tools/ncu_summary.py /tmp/r2_exact.ncu-rep $k gpurun_out/r2_ncu_$k.csv; done
ls -la gpurun_out
# launch list of two whole steps (every kernel of the process: torch's generators of the synthetic batch included;
# profiles/r2_summary.md keeps the library's own)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_all.csv python tools/prof_step.py --steps 2 --warmup 2 > gpurun_out/final_ncu_list.log 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/r2_launches_all.csv")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
keep = [rows[hi], rows[hi + 1]] + [r for r in rows[hi + 2:] if len(r) > 4 and "sdrjfm" in r[4]]
csv.writer(open("gpurun_out/r2_launches.csv", "w", newline="")).writerows(keep)
print("kept", len(keep) - 2, "launches of the library")
PY
rm -f gpurun_out/r2_launches_all.csv
