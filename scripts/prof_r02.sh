#!/bin/bash
# Round-2 profiling pass (run on the GPU box through gpurun): launch list of the bench command + `ncu --set full`
# captures of the dominant kernels; raw pages exported to CSV next to the reports.
O=gpurun_out
mkdir -p $O
NCU="ncu --set full --clock-control none --import-source on"
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/r02_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r02_bench_under_ncu.json 2> $O/r02_bench_under_ncu.err
$NCU -k regex:dft_kernel -c 1 -o $O/r02_dft_fp32 -f python scripts/prof_dft.py C3 0 1 > $O/r02_prof_dft_fp32.log 2>&1
$NCU -k regex:dft_tc5_kernel -c 1 -o $O/r02_dft_tc5 -f python scripts/prof_dft.py C3 200 1 > $O/r02_prof_dft_tc5.log 2>&1
$NCU -k regex:grid_ -s 14 -c 14 -o $O/r02_grid -f python scripts/prof_grid.py 2 > $O/r02_prof_grid.log 2>&1
$NCU -k regex:'ifft_rows|fft_chi2' -s 3 -c 3 -o $O/r02_fft -f python scripts/prof_fft.py > $O/r02_prof_fft.log 2>&1
for r in r02_dft_fp32 r02_dft_tc5 r02_grid r02_fft; do
    ncu -i $O/$r.ncu-rep --page raw --csv > $O/$r.raw.csv 2>/dev/null
done
ls -la $O | tail -20
tail -3 $O/r02_prof_grid.log $O/r02_prof_dft_tc5.log
